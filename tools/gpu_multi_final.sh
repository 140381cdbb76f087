# N-GPU lease, final sources: sharded-plane parity tests (worlds 2 / 4 / 8 as far as the lease goes), then the strong-scaling
# lines of config 4 (azimuth shards, peer stores) and config 5 (trajectory, pose shards). Usage: bash tools/gpu_multi_final.sh N TAG
cd /root/repo; N=${1:-8}; TAG=${2:-r2b}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $N "$@"; }
python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short > gpurun_out/${TAG}_multi_pytest_n$N.log 2>&1; tail -3 gpurun_out/${TAG}_multi_pytest_n$N.log
run --config 4 --shard azimuth --steps 100 --cpu-frames 0 > gpurun_out/${TAG}_bench_cfg4_az_p2p_n$N.json 2> gpurun_out/${TAG}_bench_cfg4_az_p2p_n$N.err
run --config 5 --steps 3 --warmup 1 --cpu-frames 0        > gpurun_out/${TAG}_bench_cfg5_n$N.json 2> gpurun_out/${TAG}_bench_cfg5_n$N.err
for f in gpurun_out/${TAG}_bench_cfg4_az_p2p_n$N.json gpurun_out/${TAG}_bench_cfg5_n$N.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'value %.1f e2e %.1f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d.get('single_frame_ms'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
