"""Knob sweep, round 1b (one process per config; knobs via env). Prints per-kernel ms for a 64-pair batch and single-pair us/iter."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tune import CHILD

def run(env):
    e = dict(os.environ); e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True)
    print(json.dumps(env), "->", r.stdout.strip() or r.stderr.strip()[-400:]); sys.stdout.flush()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "A"):
        for tpb, qpt, s, qb in ((512, 4, 4, 512), (512, 2, 2, 512), (512, 2, 4, 512), (512, 4, 2, 512), (1024, 2, 4, 512), (1024, 4, 4, 512), (1024, 4, 8, 512),
                                (512, 4, 8, 512), (512, 4, 4, 256), (256, 4, 2, 512), (256, 4, 4, 256), (512, 4, 4, 1024), (1024, 4, 4, 1024)):
            run({"TUNE_PAIRS": 64, "ICP_B200_TPB": tpb, "ICP_B200_QPT": qpt, "ICP_B200_S": s, "ICP_B200_QB": qb})
        run({"TUNE_PAIRS": 64, "ICP_B200_PAR_RANK": 0})
    if which in ("all", "C"):
        run({"TUNE_PAIRS": 64, "ICP_B200_CMODE": 0})
        for cc in (1, 2, 4, 8):
            for qi in (8, 16, 32):
                run({"TUNE_PAIRS": 64, "ICP_B200_CMODE": 1, "ICP_B200_CC": cc, "ICP_B200_QI": qi})
    if which in ("all", "L"):
        run({"TUNE_PAIRS": 0, "ICP_B200_CMODE": 0})
        for qi in (4, 8, 16, 32):
            run({"TUNE_PAIRS": 0, "ICP_B200_CMODE": 1, "ICP_B200_QI": qi})
        for tpb, qpt, s, qb in ((1024, 2, 16, 112), (1024, 2, 8, 112), (512, 2, 8, 112), (512, 4, 16, 112), (1024, 4, 32, 112), (512, 2, 8, 56), (256, 2, 4, 56), (1024, 2, 32, 56)):
            run({"TUNE_PAIRS": 0, "ICP_B200_TPB": tpb, "ICP_B200_QPT": qpt, "ICP_B200_S": s, "ICP_B200_QB": qb})
