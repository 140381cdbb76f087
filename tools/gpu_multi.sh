# Multi-GPU evidence on an N-GPU lease: sharded-plane parity tests, then bench lines for config 2 (pose shards + the
# azimuth-sharded evidence leg), config 4 (azimuth shards, both exchanges) and config 5 (trajectory, pose shards).
# Usage: bash tools/gpu_multi.sh N TAG [quick]
cd /root/repo; N=${1:-2}; TAG=${2:-r2}; QUICK=$3
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $N "$@"; }
python -m pytest tests/test_gpu_multi.py -m gpu -q -s --tb=short > gpurun_out/${TAG}_multi_pytest_n$N.log 2>&1; tail -4 gpurun_out/${TAG}_multi_pytest_n$N.log
run --steps 100 --cpu-frames 0                                   > gpurun_out/${TAG}_bench_cfg2_n$N.json 2> gpurun_out/${TAG}_bench_cfg2_n$N.err
run --config 4 --shard azimuth --steps 60 --cpu-frames 0         > gpurun_out/${TAG}_bench_cfg4_az_p2p_n$N.json 2> gpurun_out/${TAG}_bench_cfg4_az_p2p_n$N.err
run --config 4 --shard azimuth --exchange nccl --steps 60 --cpu-frames 0 > gpurun_out/${TAG}_bench_cfg4_az_nccl_n$N.json 2> gpurun_out/${TAG}_bench_cfg4_az_nccl_n$N.err
if [ -z "$QUICK" ] && [ $N -ge 8 ]; then      # the same lease also gives the N = 4 point of the strong-scaling curve
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus 4 --config 4 --shard azimuth --steps 60 --cpu-frames 0 > gpurun_out/${TAG}_bench_cfg4_az_p2p_n4.json 2> gpurun_out/${TAG}_bench_cfg4_az_p2p_n4.err
  python - gpurun_out/${TAG}_bench_cfg4_az_p2p_n4.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], 'value %.1f e2e %.1f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d.get('single_frame_ms'))
PY
fi
if [ -z "$QUICK" ]; then
run --config 5 --steps 2 --warmup 1 --cpu-frames 0               > gpurun_out/${TAG}_bench_cfg5_n$N.json 2> gpurun_out/${TAG}_bench_cfg5_n$N.err
run --config 2 --shard azimuth --steps 60 --cpu-frames 0         > gpurun_out/${TAG}_bench_cfg2_az_p2p_n$N.json 2> gpurun_out/${TAG}_bench_cfg2_az_p2p_n$N.err
fi
for f in gpurun_out/${TAG}_bench_*_n$N.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'value %.1f e2e %.1f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d.get('single_frame_ms'), d.get('azimuth_sharded'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
