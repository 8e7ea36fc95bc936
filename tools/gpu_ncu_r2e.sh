# ncu --set full of the mid-sized kernels of the Ed25519 table (candidates for the next optimisation): raw pages only
set -x
export TMX_SERIAL_TABLES=1
mkdir -p gpurun_out/ncu3
rm -f gpurun_out/ncu3/*
cap() {  # name regex skip count
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/ncu3/$1 python tools/profile_prove.py 1 > gpurun_out/ncu3/$1.log 2>&1
  ncu -i gpurun_out/ncu3/$1.ncu-rep --page raw --csv > gpurun_out/ncu3/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu3/$1.ncu-rep --page source --csv > gpurun_out/ncu3/$1.source.csv 2>/dev/null
  rm -f gpurun_out/ncu3/$1.ncu-rep
}
cap fri_batch fri_batch_kernel 2 1
cap eval_columns eval_columns_kernel 8 4
cap quotient_ed quotient_kernel 2 1
cap bus_count bus_count_kernel 1 1
cap bus_gen bus_gen_kernel 0 1
du -sh gpurun_out/ncu3
for f in fri_batch eval_columns quotient_ed bus_count bus_gen; do python tools/ncu_key_metrics.py gpurun_out/ncu3/$f.raw.csv > gpurun_out/ncu3/$f.summary.txt 2>&1; done
head -40 gpurun_out/ncu3/eval_columns.summary.txt
