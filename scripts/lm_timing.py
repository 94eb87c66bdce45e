"""Development: cycles per phase of the odometry solve (needs the -DLL_LM_TIMING build; LL_LIB_PATH points at it)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LL_DEBUG_LM"] = "1"
ll = importlib.import_module("light-loam_b200")
scans = [ll.synth.scan(64, k, mode=1) for k in range(12)]
ctx = ll.Context(scan_line=64, batch=1, enable_mapping=0)
for k in range(12):
    ctx.process_scans([scans[k]])
    if k >= 10:
        st = ctx.stats()
        print(k, list(st.lm_jacobian_evals), list(st.lm_cost_evals), flush=True)
