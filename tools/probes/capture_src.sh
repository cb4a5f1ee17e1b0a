# ncu captures with source correlation for per-line stall analysis; usage: capture_src.sh p0|p5|p5qt|p5f32 (<= 64 MiB come back per call)
case ${1:-p0} in
  p0) B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-outlier-leg --no-configs";;
  p5) B="python bench.py --noise 1.3 --slab-log2 28 --steps 2 --warmup 1 --no-cpu --no-e2e --no-outlier-leg --no-configs";;
  p5qt) B="python bench.py --qt --noise 1.3 --slab-log2 28 --steps 2 --warmup 1 --no-cpu --no-e2e --no-outlier-leg --no-configs";;
  p5f32) B="python bench.py --f32 --noise 1.3 --slab-log2 28 --steps 2 --warmup 1 --no-cpu --no-e2e --no-outlier-leg --no-configs";;
esac
if [ "$1" = fused ]; then  # the single-launch kernels on c1 (first field of tools/phase_times.py)
  ncu --set full --clock-control none --import-source on -k regex:"k_compress_fused|k_decompress_fused" -s 4 -c 2 -o gpurun_out/src_$1 -f python tools/phase_times.py > /dev/null 2>&1
else
  ncu --set full --clock-control none --import-source on -k regex:"k_compress|k_decompress" -s 6 -c 3 -o gpurun_out/src_$1 -f $B > /dev/null 2>&1
fi
ls -la gpurun_out/src_$1.ncu-rep
