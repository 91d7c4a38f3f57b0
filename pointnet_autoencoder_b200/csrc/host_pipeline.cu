// host_pipeline.cu -- streaming host-buffer form of the Chamfer step (NnDistance + NnDistanceGrad).
//
// The reference's TF op is fed from host memory by the session (train.py:196-206 builds the batch in numpy and
// feed_dict copies it in; the results come back through sess.run).  This is that path for a caller of the C ABI:
// every step copies its inputs host->device, runs the step's CUDA graph and copies the results device->host, with
// `depth` buffer sets on three streams so that the H2D copy of submission i+1, the kernels of submission i and the D2H
// copy of submission i-1 overlap.  A submission is `steps` consecutive batches (steps >= 1): their inputs go in with two
// copies, their kernels are one CUDA graph (programmatic dependent launch chains the steps without gaps), all their
// results come back with one copy.  One pnae_chamfer_host_pipeline_submit() call is: two cudaMemcpyAsync (inputs), one
// cudaGraphLaunch, one cudaMemcpy(2D)Async (results), three event records, two stream waits -- nothing else on the host.
// Measured on B200 (B=32, N=M=2048): one batch per submission 69 us per step, four batches per submission 54 us per
// step (the kernels alone take 52.8 us) -- a one-step graph pays its launch ramp and loses the overlap between steps.
//
// Memory stays the caller's (device inputs / outputs / workspace and the pinned result buffers are passed in at
// creation); the handle owns streams, events and graphs only.
#include <vector>

#include "pnae_common.cuh"

namespace {

struct Set {
    float *d_xyz1, *d_xyz2;
    char *d_out, *h_out;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    void *pipelined;             // pnae_chamfer_graph_create_pipelined handle (fused sets with room for two workspaces), else NULL
    cudaEvent_t ev_in, ev_run, ev_out;
    bool busy;
};

struct HostPipeline {
    int depth, steps, b, n, m;
    size_t in1_bytes, in2_bytes, d2h_bytes, out_stride;
    cudaStream_t s_in, s_run, s_out;
    std::vector<Set> sets;
    long long count;
};

void destroy(HostPipeline *hp)
{
    for (Set &s : hp->sets) {
        if (s.pipelined) pnae_graph_destroy(s.pipelined);
        if (s.exec) cudaGraphExecDestroy(s.exec);
        if (s.graph) cudaGraphDestroy(s.graph);
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_run) cudaEventDestroy(s.ev_run);
        if (s.ev_out) cudaEventDestroy(s.ev_out);
    }
    if (hp->s_in) cudaStreamDestroy(hp->s_in);
    if (hp->s_run) cudaStreamDestroy(hp->s_run);
    if (hp->s_out) cudaStreamDestroy(hp->s_out);
    delete hp;
}

}  // namespace

// One set's CUDA graph: `steps` consecutive Chamfer steps over the set's own inputs and result blocks (programmatic
// dependent launch chains them, so a graph of several steps has no gaps between steps), captured here rather than through
// pnae_chamfer_graph_create_* because every step has its own outputs.
static int capture_set(HostPipeline *hp, Set &s, int fused, const size_t *off, const float *gd1, const float *gd2,
                       void *workspace, size_t workspace_bytes)
{
    // Several fused steps and a workspace twice the size of one: the software-pipelined graph (step k+1's sweep runs while
    // step k's finalize resolves; every step has its own result block, the two workspace halves alternate).
    const size_t one = pnae_nn_distance_workspace_bytes(hp->b, hp->n, hp->m);
    const size_t half = (one + 255) & ~(size_t)255;
    if (fused && hp->steps >= 2 && workspace_bytes >= half + one) {
        std::vector<const float *> x1(hp->steps), x2(hp->steps);
        std::vector<float *> g1(hp->steps), g2(hp->steps), d1(hp->steps), d2(hp->steps);
        std::vector<int *> i1(hp->steps), i2(hp->steps);
        for (int k = 0; k < hp->steps; k++) {
            char *o = s.d_out + (size_t)k * hp->out_stride;
            x1[k] = s.d_xyz1 + (size_t)k * hp->b * hp->n * 3; x2[k] = s.d_xyz2 + (size_t)k * hp->b * hp->m * 3;
            g1[k] = (float *)(o + off[0]); g2[k] = (float *)(o + off[1]);
            d1[k] = (float *)(o + off[2]); i1[k] = (int *)(o + off[3]);
            d2[k] = (float *)(o + off[4]); i2[k] = (int *)(o + off[5]);
        }
        // three workspaces when there is room (and at least three steps): the sweeps then follow each other without a gap
        const int nws = (hp->steps >= 3 && workspace_bytes >= 2 * half + one) ? 3 : 2;
        void *ws2[3] = {workspace, (char *)workspace + half, (char *)workspace + 2 * half};
        return pnae_chamfer_graph_create_pipelined(1, hp->steps, hp->steps, nws, hp->b, hp->n, x1.data(), hp->m, x2.data(), d1.data(), i1.data(),
                                                   d2.data(), i2.data(), gd1, gd2, g1.data(), g2.data(), ws2, one, &s.pipelined);
    }
    cudaStream_t st;
    PNAE_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (ce != cudaSuccess) {
        cudaStreamDestroy(st);
        pnae_set_error("host_pipeline_create: cudaStreamBeginCapture failed: %s", cudaGetErrorString(ce));
        return PNAE_ERR_CUDA;
    }
    int rc = PNAE_OK;
    for (int k = 0; k < hp->steps && rc == PNAE_OK; k++) {
        char *o = s.d_out + (size_t)k * hp->out_stride;
        const float *x1 = s.d_xyz1 + (size_t)k * hp->b * hp->n * 3, *x2 = s.d_xyz2 + (size_t)k * hp->b * hp->m * 3;
        float *g1 = (float *)(o + off[0]), *g2 = (float *)(o + off[1]);
        float *dist1 = (float *)(o + off[2]); int *idx1 = (int *)(o + off[3]);
        float *dist2 = (float *)(o + off[4]); int *idx2 = (int *)(o + off[5]);
        if (fused) {
            rc = pnae_nn_distance_fwd_grad(hp->b, hp->n, x1, hp->m, x2, gd1, gd2, dist1, idx1, dist2, idx2, g1, g2, workspace, workspace_bytes, st);
        } else {
            rc = pnae_nn_distance_fwd(hp->b, hp->n, x1, hp->m, x2, dist1, idx1, dist2, idx2, workspace, workspace_bytes, st);
            if (rc == PNAE_OK) rc = pnae_nn_distance_bwd(hp->b, hp->n, x1, hp->m, x2, gd1, idx1, gd2, idx2, g1, g2, st);
        }
    }
    ce = cudaStreamEndCapture(st, &graph);
    cudaStreamDestroy(st);
    if (rc != PNAE_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess || graph == nullptr) {
        pnae_set_error("host_pipeline_create: cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
        return PNAE_ERR_CUDA;
    }
    ce = cudaGraphInstantiate(&s.exec, graph, 0);
    s.graph = graph;
    if (ce != cudaSuccess) {
        pnae_set_error("host_pipeline_create: cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
        return PNAE_ERR_CUDA;
    }
    return PNAE_OK;
}

extern "C" int pnae_chamfer_host_pipeline_create(int depth, int steps, int b, int n, int m, int fused,
                                                 float *const *d_xyz1, float *const *d_xyz2,
                                                 void *const *d_out, void *const *h_out, const size_t *out_offsets,
                                                 size_t out_stride, size_t d2h_bytes,
                                                 const float *grad_dist1, const float *grad_dist2,
                                                 void *workspace, size_t workspace_bytes, void **handle)
{
    PNAE_REQUIRE(handle != nullptr, "host_pipeline_create: NULL handle");
    *handle = nullptr;
    PNAE_REQUIRE(depth >= 1 && depth <= 64 && steps >= 1 && steps <= 64 && b >= 1 && n >= 1 && m >= 1, "host_pipeline_create: invalid sizes");
    PNAE_REQUIRE(d2h_bytes >= 1 && d2h_bytes <= out_stride, "host_pipeline_create: need 1 <= d2h_bytes <= out_stride");
    PNAE_REQUIRE(d_xyz1 && d_xyz2 && d_out && h_out && out_offsets && grad_dist1 && grad_dist2, "host_pipeline_create: NULL pointer");
    HostPipeline *hp = new HostPipeline();
    hp->depth = depth; hp->steps = steps; hp->b = b; hp->n = n; hp->m = m;
    hp->in1_bytes = sizeof(float) * 3 * (size_t)steps * b * n;
    hp->in2_bytes = sizeof(float) * 3 * (size_t)steps * b * m;
    hp->d2h_bytes = d2h_bytes; hp->out_stride = out_stride;
    hp->count = 0;
    hp->s_in = hp->s_run = hp->s_out = nullptr;
    hp->sets.assign(depth, Set{});
    for (Set &z : hp->sets) { z.graph = nullptr; z.exec = nullptr; z.ev_in = z.ev_run = z.ev_out = nullptr; }
    int rc = PNAE_OK;
    auto cuda_ok = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == PNAE_OK) {
            pnae_set_error("host_pipeline_create: %s failed: %s", what, cudaGetErrorString(e));
            rc = PNAE_ERR_CUDA;
        }
    };
    cuda_ok(cudaStreamCreateWithFlags(&hp->s_in, cudaStreamNonBlocking), "cudaStreamCreate");
    cuda_ok(cudaStreamCreateWithFlags(&hp->s_run, cudaStreamNonBlocking), "cudaStreamCreate");
    cuda_ok(cudaStreamCreateWithFlags(&hp->s_out, cudaStreamNonBlocking), "cudaStreamCreate");
    for (int i = 0; i < depth && rc == PNAE_OK; i++) {
        Set &s = hp->sets[i];
        s.d_xyz1 = d_xyz1[i]; s.d_xyz2 = d_xyz2[i];
        s.d_out = (char *)d_out[i]; s.h_out = (char *)h_out[i];
        s.busy = false;
        if (!s.d_xyz1 || !s.d_xyz2 || !s.d_out || !s.h_out) {
            pnae_set_error("host_pipeline_create: NULL buffer in set %d", i);
            rc = PNAE_ERR_INVALID_ARG;
            break;
        }
        cuda_ok(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming), "cudaEventCreate");
        cuda_ok(cudaEventCreateWithFlags(&s.ev_run, cudaEventDisableTiming), "cudaEventCreate");
        cuda_ok(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming), "cudaEventCreate");
        if (rc != PNAE_OK) break;
        // per-step result block (byte offsets): grad_xyz1, grad_xyz2, dist1, idx1, dist2, idx2; blocks are out_stride apart
        rc = capture_set(hp, s, fused, out_offsets, grad_dist1, grad_dist2, workspace, workspace_bytes);
    }
    if (rc != PNAE_OK) { destroy(hp); return rc; }
    *handle = hp;
    return PNAE_OK;
}

extern "C" int pnae_chamfer_host_pipeline_submit(void *handle, const float *h_xyz1, const float *h_xyz2, int *retired)
{
    PNAE_REQUIRE(handle && h_xyz1 && h_xyz2 && retired, "host_pipeline_submit: NULL pointer");
    HostPipeline *hp = static_cast<HostPipeline *>(handle);
    Set &s = hp->sets[hp->count % hp->depth];
    PNAE_REQUIRE(!s.busy, "host_pipeline_submit: buffer set still in flight");
    PNAE_CUDA_OK(cudaMemcpyAsync(s.d_xyz1, h_xyz1, hp->in1_bytes, cudaMemcpyHostToDevice, hp->s_in));
    PNAE_CUDA_OK(cudaMemcpyAsync(s.d_xyz2, h_xyz2, hp->in2_bytes, cudaMemcpyHostToDevice, hp->s_in));
    PNAE_CUDA_OK(cudaEventRecord(s.ev_in, hp->s_in));
    PNAE_CUDA_OK(cudaStreamWaitEvent(hp->s_run, s.ev_in, 0));
    if (s.pipelined) { const int rc = pnae_graph_launch(s.pipelined, hp->s_run); if (rc) return rc; }
    else PNAE_CUDA_OK(cudaGraphLaunch(s.exec, hp->s_run));
    PNAE_CUDA_OK(cudaEventRecord(s.ev_run, hp->s_run));
    PNAE_CUDA_OK(cudaStreamWaitEvent(hp->s_out, s.ev_run, 0));
    if (hp->d2h_bytes == hp->out_stride || hp->steps == 1)
        PNAE_CUDA_OK(cudaMemcpyAsync(s.h_out, s.d_out, hp->steps == 1 ? hp->d2h_bytes : hp->out_stride * hp->steps, cudaMemcpyDeviceToHost, hp->s_out));
    else          // only the leading d2h_bytes of every step's block (e.g. the gradients): one strided copy
        PNAE_CUDA_OK(cudaMemcpy2DAsync(s.h_out, hp->out_stride, s.d_out, hp->out_stride, hp->d2h_bytes, hp->steps, cudaMemcpyDeviceToHost, hp->s_out));
    PNAE_CUDA_OK(cudaEventRecord(s.ev_out, hp->s_out));
    s.busy = true;
    hp->count++;
    // free the buffer set the NEXT submit will use: its results stay valid until that submit
    Set &nx = hp->sets[hp->count % hp->depth];
    *retired = -1;
    if (nx.busy) {
        PNAE_CUDA_OK(cudaEventSynchronize(nx.ev_out));
        nx.busy = false;
        *retired = (int)(hp->count % hp->depth);
    }
    return PNAE_OK;
}

extern "C" int pnae_chamfer_host_pipeline_drain(void *handle, int *retired, int *count)
{
    PNAE_REQUIRE(handle && retired && count, "host_pipeline_drain: NULL pointer");
    HostPipeline *hp = static_cast<HostPipeline *>(handle);
    int k = 0;
    for (int i = 0; i < hp->depth; i++) {          // oldest first
        const int idx = (int)((hp->count + i) % hp->depth);
        Set &s = hp->sets[idx];
        if (!s.busy) continue;
        PNAE_CUDA_OK(cudaEventSynchronize(s.ev_out));
        s.busy = false;
        retired[k++] = idx;
    }
    *count = k;
    return PNAE_OK;
}

extern "C" int pnae_chamfer_host_pipeline_destroy(void *handle)
{
    if (handle == nullptr) return PNAE_OK;
    HostPipeline *hp = static_cast<HostPipeline *>(handle);
    cudaStreamSynchronize(hp->s_in); cudaStreamSynchronize(hp->s_run); cudaStreamSynchronize(hp->s_out);
    destroy(hp);
    return PNAE_OK;
}
