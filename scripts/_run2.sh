mkdir -p gpurun_out
t0=$(date +%s.%N); python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; t1=$(date +%s.%N); echo "default bench wall: $(echo "$t1 - $t0" | bc) s"; cat gpurun_out/bench_default.json | cut -c1-3000
t0=$(date +%s.%N); python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; t1=$(date +%s.%N); echo "reference arm wall: $(echo "$t1 - $t0" | bc) s"; cat gpurun_out/bench_reference.json
