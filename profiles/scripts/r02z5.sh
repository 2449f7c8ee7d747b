mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mdjit.py -m gpu -x -q 2>&1 | tail -5
PYTHONPATH=$PWD timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_md_tile -s 1 -c 1 -f -o gpurun_out/prof_mdgen python profiles/generic_md_profile.py > gpurun_out/ncu_mdgen.log 2>&1; echo ncu rc=$?
ncu -i gpurun_out/prof_mdgen.ncu-rep --page raw --csv > gpurun_out/mdgen_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_mdgen.ncu-rep --page source --csv > gpurun_out/mdgen_source.csv 2>/dev/null
rm -f gpurun_out/prof_mdgen.ncu-rep
ls -la gpurun_out | tail -5
