mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2w_pytest.log
for pdl in 1 0; do
CCB_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 3 --no-profile --no-cpu-baseline --no-side-configs --no-reference-gpu > gpurun_out/r2w_bench_pdl$pdl.json 2> gpurun_out/r2w_bench_pdl$pdl.err; echo "bench pdl=$pdl rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2w_bench_pdl$pdl.json'));print('pdl=$pdl', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'loss', d['loss'])"; tail -n 2 gpurun_out/r2w_bench_pdl$pdl.err | cut -c1-300
done
