#!/bin/bash
# round-2 GPU call 17: A/B lean4 (product) / lean5 (+ L2 evict-first on the streamed row blocks) / lean6 (+ panel warp loads its
# fragments before the barrier); DRAM bytes of the isolated launch for each
mkdir -p gpurun_out
timeout 300 python tools/qr_variants.py tools/_variants/lean4.so tools/_variants/lean5.so tools/_variants/lean6.so > gpurun_out/c17_qr_ab.log 2>&1
cat gpurun_out/c17_qr_ab.log
for v in lean4 lean5 lean6; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_test_qr_ft --launch-skip 1 -c 1 --csv --log-file gpurun_out/c17_dram_$v.csv python tools/qr_one_lib.py tools/_variants/$v.so > /dev/null 2>&1
  tail -4 gpurun_out/c17_dram_$v.csv | cut -d, -f10-
done
