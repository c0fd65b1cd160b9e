mkdir -p gpurun_out
T=${1:-r3x}
rm -f gpurun_out/${T}_* gpurun_out/r02_parity.txt
( timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -s 2>&1 | grep -E "FULLSIZE|passed|failed|FAILED|Error" ) > gpurun_out/${T}_test_all.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/${T}_bench.err | tail -1 ) > gpurun_out/${T}_bench_n1.json
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1
