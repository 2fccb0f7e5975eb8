python -m pytest tests/test_of_gpu.py tests/test_golden.py tests/test_ref_fixtures_gpu.py -q -m gpu 2>&1 | tail -3
python tools/bench_of.py --only grc > gpurun_out/of_grc_f2d.json 2> gpurun_out/of_grc_f2d.err
grep -h "ms_per_rollout\|finite" gpurun_out/of_grc_f2d.json
