# Round 2, re-entry (1 GPU), the safe part: what host memory this box gives a job, the kernel-variant A/B, BASELINE configs[3] to
# depth 13 entirely in HBM.
mkdir -p gpurun_out
( free -g | head -2; nproc; for f in /sys/fs/cgroup/memory.max /sys/fs/cgroup/memory.current /sys/fs/cgroup/memory/memory.limit_in_bytes; do [ -r $f ] && echo "$f $(cat $f)"; done; ulimit -l ) 2>&1 | tee gpurun_out/host_memory.txt
python -c "import bench; print('bench.host_memory_available GB', bench.host_memory_available() / 1e9)" | tee -a gpurun_out/host_memory.txt
bash tools/gpu_r2b_ab.sh
python -c "
import _pkg; pkg=_pkg.load(); open('gpurun_out/cfg4.cfg','w').write(pkg.cfg_text(5, ['v1','v2'], 2))"
echo "== cfg4 depth 13, no spill"
( time timeout 300 vsr-tlaplus_b200/vsrmc -deadlock -notrace -depth 13 -table 1300000000 -frontier 720000000 -config gpurun_out/cfg4.cfg ) 2>&1 | tail -22 | tee gpurun_out/cfg4_depth13.txt
