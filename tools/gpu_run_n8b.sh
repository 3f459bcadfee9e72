# 8-GPU bench line only (value, e2e, sequence pipeline with the NCCL gather)
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
d=json.loads(open('gpurun_out/bench_n%s.json' % N).read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'parity', d.get('parity_checked',{}).get('ok'))
for k, v in d.get('pipeline', {}).items(): print(' pipeline', k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ('frames_per_s','ms_per_frame_step','host_pnp_post_ms_per_frame_step_max_rank','pnp_workers_per_rank','gather_ms','poses_solved_frac')})
PY
