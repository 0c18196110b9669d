#!/bin/bash
# Last check of a round on ONE B200: smoke(), the whole GPU suite, the default bench line, the launch list of the c2 step.
P=${1:-r2z}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2 > gpurun_out/${P}_smoke.txt; cat gpurun_out/${P}_smoke.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${P}_pytest.txt; cat gpurun_out/${P}_pytest.txt
timeout 400 python bench.py > gpurun_out/${P}_c2_default.json 2> gpurun_out/${P}_c2_default.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 62 -c 64 --csv --log-file gpurun_out/${P}_launches_c2_b4096.csv python tools/prof_step.py --batch 4096 --steps 4 --gemm-mode 1 > gpurun_out/${P}_prof.log 2>&1
tail -1 gpurun_out/${P}_prof.log
