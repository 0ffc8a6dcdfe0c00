"""A/B the K1 variants inside one process environment each (same box)."""
import os, subprocess, sys
for var in (4, 3, 4, 3):
    r = subprocess.run([sys.executable, "tools/bench_k1.py"], env=dict(os.environ, COCOS_K1_VARIANT=str(var)), capture_output=True, text=True)
    import json
    try:
        d = json.loads(r.stdout[r.stdout.index("{"):])
        print("variant", var, {k: round(v["ms"], 4) for k, v in d.items() if k.startswith("k1_fwd")}, flush=True)
    except Exception as e:
        print("variant", var, "failed", r.stdout[-300:], r.stderr[-500:])
