"""One lane, synchronous calls: the process profiled by scripts/gpu_latlist.sh (launch list of the single-stream path)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ll = importlib.import_module("light-loam_b200")
scans = [ll.synth.scan(64, k, mode=1) for k in range(14)]
ctx = ll.Context(scan_line=64, batch=1, enable_mapping=0)
for k in range(14):
    ctx.process_scans([scans[k]])
print(ctx.stats().kernel_launches)
