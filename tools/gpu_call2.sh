mkdir -p gpurun_out
echo "nproc=$(nproc) cpu.max=$(cat /sys/fs/cgroup/cpu.max 2>/dev/null) cfs=$(cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us 2>/dev/null)/$(cat /sys/fs/cgroup/cpu/cpu.cfs_period_us 2>/dev/null)" > gpurun_out/cpuinfo.txt
python -c "import os; print('affinity', len(os.sched_getaffinity(0)), 'cpu_count', os.cpu_count())" >> gpurun_out/cpuinfo.txt
grep -c processor /proc/cpuinfo >> gpurun_out/cpuinfo.txt; free -g | head -2 >> gpurun_out/cpuinfo.txt; nvidia-smi -L >> gpurun_out/cpuinfo.txt
cat gpurun_out/cpuinfo.txt
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/t_all2.log 2>&1; echo "all rc=$?"; tail -15 gpurun_out/t_all2.log
