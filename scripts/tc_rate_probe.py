"""tcgen05.mma kind::tf32 layout / issue-rate table from the standalone probe (tests/probes/tc_probe.cu).

    python scripts/tc_rate_probe.py [out.json]

Every configuration runs in its own process (a faulting descriptor only kills that run).  Modes: 0 K-major no-swizzle,
1 MN-major no-swizzle, 2 K-major SWIZZLE_128B, 3 (A) tensor memory."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "probes", "_bin", "tc_probe")


def run(N, K, passes, reps, a_mode, b_mode, a_lbo=128, mn_sbo=128, swap_mn=0):
    try:
        out = subprocess.run([BIN] + [str(x) for x in (N, K, passes, reps, a_mode, b_mode, a_lbo, mn_sbo, swap_mn)], capture_output=True,
                             text=True, timeout=60)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else ""
        r = json.loads(line) if line.startswith("{") else {"error": (out.stderr or "no output").strip()[-200:]}
    except subprocess.TimeoutExpired:
        r = {"error": "timeout"}
    r.update(dict(N=N, K=K, passes=passes, reps=reps, a_mode=a_mode, b_mode=b_mode, a_lbo=a_lbo, mn_sbo=mn_sbo, swap_mn=swap_mn))
    return r


def main():
    rows = []
    # layout correctness (reps = 1) -------------------------------------------------------------------
    for a_mode, b_mode, a_lbo, mn_sbo, swap in [(0, 0, 128, 128, 0), (0, 0, 144, 128, 0), (3, 0, 128, 128, 0), (2, 2, 128, 128, 0), (2, 0, 128, 128, 0),
                                                (0, 2, 128, 128, 0),
                                                (0, 1, 128, 128, 0), (0, 1, 128, 128, 1), (0, 1, 128, 144, 0), (0, 1, 128, 144, 1),
                                                (1, 0, 128, 128, 0), (1, 0, 128, 128, 1), (1, 1, 128, 144, 0), (1, 1, 128, 144, 1), (3, 1, 128, 144, 0),
                                                (3, 1, 128, 144, 1)]:
        for N, K in ((112, 32), (256, 64), (128, 128)):
            rows.append(dict(kind="layout", **run(N, K, 3, 1, a_mode, b_mode, a_lbo, mn_sbo, swap)))
            print(json.dumps(rows[-1]), flush=True)
    # issue rate (reps = 400, passes = 1) -----------------------------------------------------------------
    for a_mode, b_mode, a_lbo, mn_sbo, swap in [(0, 0, 128, 128, 0), (0, 0, 144, 128, 0), (2, 2, 128, 128, 0), (2, 0, 128, 128, 0), (3, 0, 128, 128, 0),
                                                (3, 2, 128, 128, 0), (0, 1, 128, 144, 0), (0, 1, 128, 144, 1), (1, 1, 128, 144, 0), (1, 1, 128, 144, 1),
                                                (3, 1, 128, 144, 0), (3, 1, 128, 144, 1), (1, 0, 128, 128, 0), (1, 0, 128, 128, 1)]:
        for N in (32, 64, 112, 128, 256):
            rows.append(dict(kind="rate", **run(N, 64, 1, 400, a_mode, b_mode, a_lbo, mn_sbo, swap)))
            print(json.dumps(rows[-1]), flush=True)
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=0)


if __name__ == "__main__":
    main()
