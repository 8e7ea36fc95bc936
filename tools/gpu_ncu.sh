# ncu --set full captures of the dominant kernels of one N=128 skip proof (Ed25519 table instances); CSV pages are
# exported on the box because the merged gpurun_out/ is capped at 64 MiB.
set -x
export TMX_SERIAL_TABLES=1  # one stream, tables in order: launch indices below are deterministic
mkdir -p gpurun_out/ncu
rm -f gpurun_out/ncu/*
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/ncu/$1 python tools/profile_prove.py 1 > gpurun_out/ncu/$1.log 2>&1
  ncu -i gpurun_out/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$1.ncu-rep --page source --csv > gpurun_out/ncu/$1.source.csv 2>/dev/null
  if [ $(stat -c %s gpurun_out/ncu/$1.ncu-rep) -gt 9000000 ]; then rm gpurun_out/ncu/$1.ncu-rep; fi
}
cap leaf_hash leaf_hash_kernel 10 1
cap ntt ntt_pass_kernel 24 6
cap quotient_ed quotient_kernel 2 1   # third quotient launch = Ed25519 table
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_2proofs.csv python tools/profile_prove.py 2 > gpurun_out/launches.log 2>&1
du -sh gpurun_out
