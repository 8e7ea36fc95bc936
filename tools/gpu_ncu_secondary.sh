# ncu --set full captures (raw CSV page only) of the kernels besides K1 / K2 leaf hashing / the Ed25519 quotient, which
# tools/gpu_ncu.sh covers: one N=128 skip proof, final code of the round.  Output: gpurun_out/ncu2/*.raw.csv
set -x
export TMX_SERIAL_TABLES=1  # one stream, tables in order: launch indices below are deterministic
mkdir -p gpurun_out/ncu2
rm -f gpurun_out/ncu2/*
cap() {  # name regex skip count
  timeout 300 ncu --set full --clock-control none -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/ncu2/$1 python tools/profile_prove.py 1 > gpurun_out/ncu2/$1.log 2>&1
  ncu -i gpurun_out/ncu2/$1.ncu-rep --page raw --csv > gpurun_out/ncu2/$1.raw.csv 2>/dev/null
  rm -f gpurun_out/ncu2/$1.ncu-rep
}
cap fri_batch fri_batch_kernel 2 1      # Ed25519 table
cap eval_columns eval_columns_kernel 8 4  # Ed25519 table: constant / first-round / second-round columns, quotient
cap fri_fold fri_fold_kernel 0 12
cap ladder ed25519_ladder_kernel 0 1
cap expand ed25519_expand_kernel 0 1
cap quotient_sha quotient_kernel 0 2
cap bus_count bus_count_kernel 1 1     # round 1 order: SHA-256, Ed25519, SHA-512, logic
cap bus_gen bus_gen_kernel 0 1         # round 2 order: Ed25519, SHA-512, SHA-256, logic, range
cap bus_gen_logic bus_gen_kernel 3 1
cap merkle_levels merkle_level_kernel 0 14
cap sha256_witness 'sha256_.*_kernel' 0 4
du -sh gpurun_out/ncu2
