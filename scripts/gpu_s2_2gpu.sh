# Session-2 sanity on 2 GPUs: the driver's command line (C3, device-resident time loop, e2e leg) and the 2-D kernels in the
# same loop (they consume dt in their prologue, a one-warp kernel publishes) -- every line must say multi_gpu_bitwise true
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 "$@" > gpurun_out/s2_2gpu_$name.json 2> gpurun_out/s2_2gpu_$name.err
  echo "$name rc=$?"
  python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/s2_2gpu_$name.json') if l.startswith('{')][-1]
print('$name', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'value %.4e' % d['value'], 'frac/GPU %.3f' % d['roofline']['frac'], 'bitwise', d.get('multi_gpu_bitwise'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))"
}
run c3_default --steps 20 --warmup 3
run c2_loop --workload c2 --steps 50 --warmup 5 --no-cpu --no-e2e
run c4_unknowns_loop --workload c4 --output unknowns --steps 50 --warmup 5 --no-cpu --no-e2e
