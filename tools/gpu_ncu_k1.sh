# ncu --set full of the six ntt_pass_kernel launches of the Ed25519 table's first-round LDE (982 x 2^15) inside one proof, and of
# one leaf_hash_kernel launch of the same table; raw CSV pages are exported on the box.
set -x
mkdir -p gpurun_out/ncu
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/ncu/$1 python tools/profile_prove.py 1 > gpurun_out/ncu/$1.log 2>&1
  ncu -i gpurun_out/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1.raw.csv 2>/dev/null
  rm -f gpurun_out/ncu/$1.ncu-rep
}
# launches before the Ed25519 table's LDE: 4 constant-column batches at circuit build (6 each; the range table's 2^16 rows take
# the same plan), then the SHA-256 and SHA-512 tables of the proof
cap ntt ntt_pass_kernel 0 60
cap leaf_hash leaf_hash_kernel 0 16
python tools/ncu_key_metrics.py gpurun_out/ncu/ntt.raw.csv | grep -E "^==|dram__bytes|gpu__time_duration" > gpurun_out/ncu/ntt_summary.txt
python tools/ncu_key_metrics.py gpurun_out/ncu/leaf_hash.raw.csv | grep -E "^==|gpu__time_duration|smsp__inst_executed.sum|fmaheavy|pipe_alu|issue_active" > gpurun_out/ncu/leaf_summary.txt
tail -5 gpurun_out/ncu/ntt.log
