#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/s_bench_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/s_bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s_bench_driver.json 2> gpurun_out/s_bench_driver.err; echo "driver-style rc=$?"
timeout 900 python bench.py > gpurun_out/s_bench_default.json 2> gpurun_out/s_bench_default.err; echo "default rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/s_bench_driver.json','gpurun_out/s_bench_default.json'):
    d=json.load(open(f))
    print(f, json.dumps({k:d[k] for k in ('value','ms_per_step','steps','windows')}), 'e2e %.0f grads %.0f'%(d['e2e']['value'], d['e2e']['gradients_only']['value']),
          'frac %.4f step %.4f three %.4f'%(d['roofline']['frac'], d['roofline']['fwd_grad_step_frac'], d['roofline']['fwd_grad_step_frac_three_kernel_form']), 'cpu', d.get('cpu_baseline',{}).get('value'))
    print('  window_ms', [round(x,3) for x in d['window_ms']])
    ex=d['extra']
    print('  emd', {k:(round(v,4) if isinstance(v,float) else v) for k,v in ex['emd_strong_scaling'].items() if k in ('approx_match_ms','match_cost_fwd_grad_ms','emd_fwd_grad_ms','emd_frac_of_fp32_peak_all_gpus')}, 'mufu frac %.3f'%ex['emd_strong_scaling']['roofline']['frac'])
    print('  train', {k:round(v['samples_per_s']) for k,v in ex['ae_train'].items() if isinstance(v,dict)}, 'ref_gpu', {k:round(v,3) for k,v in ex['reference_gpu'].items() if isinstance(v,float)}, 'enc', ex['encoder_conv_pool'])
PY
timeout 300 python tools/graph_time.py --enc
timeout 300 python train_bench.py --model upconv --steps 50 | cut -c1-200
timeout 300 python train_bench.py --model upconv --steps 50 --device-pipeline | cut -c1-200
timeout 300 python train_bench.py --model emd --steps 30 | cut -c1-200
timeout 300 python train_bench.py --model fc --steps 50 | cut -c1-200
