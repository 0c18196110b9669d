#!/bin/bash
# Round-end evidence run on ONE B200 (gpurun): GPU tests, the bench lines of every config, ncu artefacts (exported to CSV on
# the box: gpurun copies back at most 64 MiB).  Outputs under gpurun_out/ with the prefix $1 (default r2f); $2 = "full" runs
# the whole GPU suite instead of the quick subset.
P=${1:-r2f}
mkdir -p gpurun_out
if [ "$2" = "full" ]; then T="tests"; else T="tests/test_gpu_first_slice.py tests/test_gpu_gru.py tests/test_gpu_golden.py tests/test_gpu_prme_k.py"; fi
timeout 900 python -m pytest $T -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${P}_pytest.txt; cat gpurun_out/${P}_pytest.txt
grep -q "failed\|error" gpurun_out/${P}_pytest.txt && exit 1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${P}_c2.json 2> gpurun_out/${P}_c2.err
timeout 300 python bench.py --config c3 --steps 20 --warmup 5 > gpurun_out/${P}_c3.json 2> gpurun_out/${P}_c3.err
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/${P}_c4.json 2> gpurun_out/${P}_c4.err
timeout 400 python bench.py --config c5 --scaling strong --batch 8192 --steps 5 --warmup 2 --no-cpu-baseline --sustain-s 0 > gpurun_out/${P}_c5_b8192_n1.json 2> gpurun_out/${P}_c5_b8192_n1.err
# ncu: launch list of two c2 steps, then full captures of the round's new / changed kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 74 -c 76 --csv --log-file gpurun_out/${P}_launches_c2_b4096.csv python tools/prof_step.py --batch 4096 --steps 4 --gemm-mode 1 > gpurun_out/${P}_prof1.log 2>&1
export_rep() {   # $1 = report stem: raw page (all metrics) + per-instruction source page, then drop the report
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | cut -d, -f1-8 > gpurun_out/$1_source.csv
  rm -f gpurun_out/$1.ncu-rep
}
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_sort_seg_fused|k_gemm_tn_tc_persist|k_rows_update_warp" -s 8 -c 8 -o gpurun_out/${P}_ncu_c2 python tools/prof_step.py --batch 4096 --steps 3 --gemm-mode 1 > gpurun_out/${P}_prof2.log 2>&1
export_rep ${P}_ncu_c2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_prme_score_warp|k_prme_apply" -s 2 -c 2 -o gpurun_out/${P}_ncu_c3 python tools/prof_mf.py --config c3 > gpurun_out/${P}_prof3.log 2>&1
export_rep ${P}_ncu_c3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"^k_gemm_tn_tc$" -s 508 -c 4 -o gpurun_out/${P}_ncu_c5 python bench.py --config c5 --scaling strong --batch 1024 --steps 1 --warmup 1 --no-parity --no-cpu-baseline --sustain-s 0 > gpurun_out/${P}_prof4.log 2>&1
export_rep ${P}_ncu_c5
du -sh gpurun_out; ls -la gpurun_out/${P}_* | awk '{print $5, $9}'
