mkdir -p gpurun_out
export_rep() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | cut -d, -f1-8 > gpurun_out/$1_source.csv
  rm -f gpurun_out/$1.ncu-rep
}
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_geoie_batch_k" -s 1 -c 1 -o gpurun_out/r2x_ncu_c4 python tools/prof_mf.py --config c4 > gpurun_out/r2x_prof_c4.log 2>&1
export_rep r2x_ncu_c4
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_rows_update_warp|k_sort_seg_fused|k_gather_rows" -c 4 -o gpurun_out/r2x_ncu_micro python tools/prof_micro.py > gpurun_out/r2x_prof_micro.log 2>&1
export_rep r2x_ncu_micro
ls -la gpurun_out/r2x_* | awk '{print $5, $9}'
