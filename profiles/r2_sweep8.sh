set -x
for c in awgn good moderate poor; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 sweep.py --frames 100000 --condition $c > gpurun_out/r2_sweep_${c}_100k_8gpu.jsonl 2> gpurun_out/r2_sweep_${c}_8gpu.err || echo "sweep $c failed"
  tail -c 400 gpurun_out/r2_sweep_${c}_100k_8gpu.jsonl
done
