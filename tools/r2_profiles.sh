#!/bin/bash
# Round-2 ncu evidence (run on the GPU box through gpurun; outputs under gpurun_out/).
set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
# 1. launch list of the headline bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_disk1m.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --no-parity > gpurun_out/r2_ncu_bench.log 2>&1
# 2. full capture of the headline kernel (slab kernel with the tail split)
ncu --set full --clock-control none --import-source on -k regex:allpairs_sym_kernel -s 1 -c 1 -f -o gpurun_out/r2_gravity_sym1m \
    python tools/profile_run.py gravity1m > gpurun_out/r2_ncu_sym1m.log 2>&1
ncu -i gpurun_out/r2_gravity_sym1m.ncu-rep --page raw --csv > gpurun_out/r2_gravity_sym1m_raw.csv
python tools/ncu_summary.py gpurun_out/r2_gravity_sym1m_raw.csv > gpurun_out/r2_gravity_sym1m_ncu_full_summary.txt
# 3. the slab-free kernel as shipped, the 64k workload, the 4M Kepler tick, the candidate passes: time + DRAM bytes
for mode in gravity1m_lean gravity64k_sym kepler4m collision256k moderef1m; do
  ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/r2_dram_$mode.csv \
      python tools/profile_run.py $mode > gpurun_out/r2_ncu_$mode.log 2>&1
done
# 4. full capture of the 4M-object vessel tick
ncu --set full --clock-control none --import-source on -k regex:vessel_tick_kernel -s 1 -c 1 -f -o gpurun_out/r2_kepler4m \
    python tools/profile_run.py kepler4m > gpurun_out/r2_ncu_kepler4m_full.log 2>&1
ncu -i gpurun_out/r2_kepler4m.ncu-rep --page raw --csv > gpurun_out/r2_kepler4m_raw.csv
python tools/ncu_summary.py gpurun_out/r2_kepler4m_raw.csv > gpurun_out/r2_kepler4m_ncu_full_summary.txt
# 5. full capture of the symmetric kernel at 64k
ncu --set full --clock-control none --import-source on -k regex:allpairs_sym_kernel -s 1 -c 1 -f -o gpurun_out/r2_gravity_sym64k \
    python tools/profile_run.py gravity64k_sym > gpurun_out/r2_ncu_sym64k.log 2>&1
ncu -i gpurun_out/r2_gravity_sym64k.ncu-rep --page raw --csv > gpurun_out/r2_gravity_sym64k_raw.csv
python tools/ncu_summary.py gpurun_out/r2_gravity_sym64k_raw.csv > gpurun_out/r2_gravity_sym64k_ncu_full_summary.txt
ls -la gpurun_out/ | tail -30
