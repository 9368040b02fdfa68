for c in 1 4 8 64 1 8; do
FLIPGPU_WAVES=$c python scripts/phase_probe.py 4096 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('waves $c', d['ms'], d['total_ms'], d['state_sha256'])" | tee -a gpurun_out/waves.txt
done
ORDER=rb FLIPGPU_WAVES=1 python scripts/phase_probe.py 4096 2>&1 | tail -1
ORDER=rb FLIPGPU_WAVES=8 python scripts/phase_probe.py 4096 2>&1 | tail -1
