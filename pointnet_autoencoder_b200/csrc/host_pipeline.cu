// host_pipeline.cu -- streaming host-buffer form of the Chamfer step (NnDistance + NnDistanceGrad).
//
// The reference's TF op is fed from host memory by the session (train.py:196-206 builds the batch in numpy and
// feed_dict copies it in; the results come back through sess.run).  This is that path for a caller of the C ABI:
// every step copies its inputs host->device, runs the step's CUDA graph and copies the results device->host, with
// `depth` buffer sets on three streams so that the H2D copy of step i+1, the kernels of step i and the D2H copy of
// step i-1 overlap.  One pnae_chamfer_host_pipeline_submit() call is: two cudaMemcpyAsync (inputs), one
// cudaGraphLaunch, one cudaMemcpyAsync (results), three event records, two stream waits -- nothing else on the host.
//
// Memory stays the caller's (device inputs / outputs / workspace and the pinned result buffers are passed in at
// creation); the handle owns streams, events and graphs only.
#include <vector>

#include "pnae_common.cuh"

namespace {

struct Set {
    float *d_xyz1, *d_xyz2;
    char *d_out, *h_out;
    void *graph;                   // pnae graph handle
    cudaEvent_t ev_in, ev_run, ev_out;
    bool busy;
};

struct HostPipeline {
    int depth, b, n, m;
    size_t in1_bytes, in2_bytes, d2h_bytes;
    cudaStream_t s_in, s_run, s_out;
    std::vector<Set> sets;
    long long count;
};

void destroy(HostPipeline *hp)
{
    for (Set &s : hp->sets) {
        if (s.graph) pnae_graph_destroy(s.graph);
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

extern "C" int pnae_chamfer_host_pipeline_create(int depth, int b, int n, int m, int fused,
                                                 float *const *d_xyz1, float *const *d_xyz2,
                                                 void *const *d_out, void *const *h_out, const size_t *out_offsets, size_t d2h_bytes,
                                                 const float *grad_dist1, const float *grad_dist2,
                                                 void *workspace, size_t workspace_bytes, void **handle)
{
    PNAE_REQUIRE(handle != nullptr, "host_pipeline_create: NULL handle");
    *handle = nullptr;
    PNAE_REQUIRE(depth >= 1 && depth <= 64 && b >= 1 && n >= 1 && m >= 1, "host_pipeline_create: invalid sizes");
    PNAE_REQUIRE(d_xyz1 && d_xyz2 && d_out && h_out && out_offsets && grad_dist1 && grad_dist2, "host_pipeline_create: NULL pointer");
    HostPipeline *hp = new HostPipeline();
    hp->depth = depth; hp->b = b; hp->n = n; hp->m = m;
    hp->in1_bytes = sizeof(float) * 3 * (size_t)b * n;
    hp->in2_bytes = sizeof(float) * 3 * (size_t)b * m;
    hp->d2h_bytes = d2h_bytes;
    hp->count = 0;
    hp->s_in = hp->s_run = hp->s_out = nullptr;
    hp->sets.assign(depth, Set{});
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
        // flat result layout (byte offsets): grad_xyz1, grad_xyz2, dist1, idx1, dist2, idx2
        float *g1 = (float *)(s.d_out + out_offsets[0]), *g2 = (float *)(s.d_out + out_offsets[1]);
        float *dist1 = (float *)(s.d_out + out_offsets[2]); int *idx1 = (int *)(s.d_out + out_offsets[3]);
        float *dist2 = (float *)(s.d_out + out_offsets[4]); int *idx2 = (int *)(s.d_out + out_offsets[5]);
        const float *x1 = s.d_xyz1, *x2 = s.d_xyz2;
        rc = fused ? pnae_chamfer_graph_create_fused_multi(1, b, n, &x1, m, &x2, dist1, idx1, dist2, idx2, grad_dist1, grad_dist2, g1, g2,
                                                           workspace, workspace_bytes, &s.graph)
                   : pnae_chamfer_graph_create_multi(1, b, n, &x1, m, &x2, dist1, idx1, dist2, idx2, grad_dist1, grad_dist2, g1, g2,
                                                     workspace, workspace_bytes, &s.graph);
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
    int rc = pnae_graph_launch(s.graph, hp->s_run);
    if (rc != PNAE_OK) return rc;
    PNAE_CUDA_OK(cudaEventRecord(s.ev_run, hp->s_run));
    PNAE_CUDA_OK(cudaStreamWaitEvent(hp->s_out, s.ev_run, 0));
    PNAE_CUDA_OK(cudaMemcpyAsync(s.h_out, s.d_out, hp->d2h_bytes, cudaMemcpyDeviceToHost, hp->s_out));
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
