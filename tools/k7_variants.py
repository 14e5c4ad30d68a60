"""Times k_normal_eq_stream builds (tools/build_variant.sh) on 20 M residual blocks: k7_variants.py name [name ...]; each name is run
in a child process with PFILTER_B200_LIB pointing at build_variants/<name>/ ("main" = the in-tree build)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
    from pf_loader import pfb
    import sweep
    n = int(os.environ.get("K7_BLOCKS", "10000000"))
    e, s = sweep.residual_blocks(n)
    H, g, c, ms = pfb.capi.eval_normal_eq_timed(sweep.SWEEP_POSE, e, s, reps=5)
    b = 128.0 * n
    print(json.dumps({"lib": os.environ.get("PFILTER_B200_LIB", "main"), "blocks": 2 * n, "ms": ms, "gbs": b / ms / 1e6, "frac": b / ms / 1e6 / 6533.5, "cost": c}))
else:
    for name in sys.argv[1:]:
        env = dict(os.environ)
        if name != "main":
            env["PFILTER_B200_LIB"] = os.path.join(ROOT, "build_variants", name, "libpfilter_b200.so")
        subprocess.run([sys.executable, __file__, "--child"], env=env)
