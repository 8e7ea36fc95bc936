# ncu --set full of the leaf hashing of the Ed25519 table's first-round LDE (982 columns x 2^16 rows: 123 permutations per leaf)
set -x
export TMX_SERIAL_TABLES=1  # one stream, tables in order SHA-256, Ed25519, SHA-512, logic, range (after 5 constant-column trees)
mkdir -p gpurun_out/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:leaf_hash_kernel -s 6 -c 1 -f -o gpurun_out/ncu/leaf_hash python tools/profile_prove.py 1 > gpurun_out/ncu/leaf_hash.log 2>&1
ncu -i gpurun_out/ncu/leaf_hash.ncu-rep --page raw --csv > gpurun_out/ncu/leaf_hash.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu/leaf_hash.ncu-rep --page source --csv > gpurun_out/ncu/leaf_hash.source.csv 2>/dev/null
rm -f gpurun_out/ncu/leaf_hash.ncu-rep
python tools/ncu_key_metrics.py gpurun_out/ncu/leaf_hash.raw.csv
python tools/ncu_source_ops.py gpurun_out/ncu/leaf_hash.source.csv 2>/dev/null | head -40
