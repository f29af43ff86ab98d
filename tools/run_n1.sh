# one-GPU validation + measurement pass of a round (run through gpurun): tests, smoke, bench lines, launch list, full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_final.log 2>&1; tail -3 gpurun_out/r02_pytest_final.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --workload fixtures --steps 5 --warmup 3 > gpurun_out/r02_fixtures_n1.json 2> gpurun_out/r02_fixtures_n1.err; echo "fixtures rc=$?"
timeout 600 python bench.py --workload chain20 --steps 5 --warmup 3 > gpurun_out/r02_chain20_n1.json 2> gpurun_out/r02_chain20_n1.err; echo "chain20 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense > gpurun_out/ncu_q34.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_expand_low' -s 3 -c 1 -f -o gpurun_out/r02_prof_low python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
