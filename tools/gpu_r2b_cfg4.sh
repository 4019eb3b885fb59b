# Round 2 (1 GPU): BASELINE configs[3] = ReplicaCount=5, Values={v1,v2}, StartViewOnTimerLimit=2 (80-byte states, 16 successors per state).
# (a) depth 13 entirely in HBM; (b) depth 14 with the frontier spilling into pinned host memory (the level alone is ~3.2e9 states = 256 GB).
mkdir -p gpurun_out
free -g | head -2
python -c "
import _pkg; pkg=_pkg.load(); open('gpurun_out/cfg4.cfg','w').write(pkg.cfg_text(5, ['v1','v2'], 2))"
echo "== depth 13, no spill"
( time timeout 300 vsr-tlaplus_b200/vsrmc -deadlock -notrace -depth 13 -table 1300000000 -frontier 720000000 -config gpurun_out/cfg4.cfg ) 2>&1 | tail -22 | tee gpurun_out/cfg4_depth13.txt
echo "== depth 14, frontier spill to pinned host memory"
( time timeout 900 vsr-tlaplus_b200/vsrmc -deadlock -notrace -depth 14 -table 5200000000 -frontier 480000000 -spill 2900000000 -config gpurun_out/cfg4.cfg ) 2>&1 | tail -24 | tee gpurun_out/cfg4_depth14_spill.txt
