# ncu --set full of one launch per mask policy of tools/mask_variants.py; summaries only (the reports stay on the box)
for v in "$@"; do
  n=$(echo "$v" | tr -d " +,()")
  G2P_ONLY="$v" python tools/mask_variants.py > gpurun_out/plain_$n.log 2>&1 && \
  G2P_ONLY="$v" ncu --set full --clock-control none --import-source on -k regex:apply_window -s 4 -c 1 -o /tmp/r2_$n -f python tools/mask_variants.py > gpurun_out/ncu_$n.log 2>&1
  python tools/ncu_summary.py rep /tmp/r2_$n.ncu-rep > gpurun_out/r2b_$n.ncu.txt
  ncu -i /tmp/r2_$n.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2b_$n.source.csv.gz
done
