mkdir -p gpurun_out
python profiles/auction_workload.py 1000
python profiles/auction_workload.py 300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_auction -c 1 -f -o gpurun_out/prof_k_auction python profiles/auction_workload.py 1000 > gpurun_out/ncu_auction.log 2>&1; echo ncu rc=$?
