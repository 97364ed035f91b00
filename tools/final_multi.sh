# usage: bash tools/final_multi.sh N  (on the GPU box; N GPUs visible)
N=$1
mkdir -p gpurun_out/final
run() { # name, extra args
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/final/bench_${name}_n$N.json 2> gpurun_out/final/bench_${name}_n$N.err
  tail -c 300 gpurun_out/final/bench_${name}_n$N.json; echo
}
run uniform
if [ "$N" = "8" ]; then
  run text --workload text_1GiB_word32
  run blocks --workload blocks_64KiB_word32
  run alias --workload zipf1.1_1GiB_alias32
fi
