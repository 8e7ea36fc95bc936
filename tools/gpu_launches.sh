# per-launch device times of two N=128 skip proofs (ncu launch list; cold-cache, serialised: shares, not absolutes)
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_2proofs.csv python tools/profile_prove.py 2 > gpurun_out/launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches_2proofs.csv > gpurun_out/launch_shares.txt 2>&1
cat gpurun_out/launch_shares.txt
