"""Generates tests/golden/oracle_golden.npz from the CPU oracle (oracle/ps_oracle.py).

The reference's arithmetic (powerbox) cannot be imported in the build container, so these
vectors freeze the *oracle's* output (parity unpinned, see oracle/ps_oracle.py header); they
guard the oracle against drift and give the GPU tests committed expected values.

Run from the repo root:  python -m tests.golden.make_golden
Inputs are produced by an explicit integer hash (no RNG-version dependence).
"""
import os

import numpy as np

CASES = {
    # name: shape, box lengths, bins, log, (ignore_kperp_zero, ignore_kpar_zero, ignore_k_zero), edge rule
    "lc_pow2_log": dict(shape=(32, 32, 64), L=(48.0, 48.0, 96.0), nb=12, log=True, flags=(True, False, False), rule="legacy_1.001"),
    "lc_pow2_lin_dc": dict(shape=(16, 16, 32), L=(24.0, 24.0, 48.0), nb=8, log=False, flags=(False, False, False), rule="legacy_1.001"),
    "coeval_200ish": dict(shape=(20, 20, 20), L=(30.0, 30.0, 30.0), nb=9, log=True, flags=(True, False, False), rule="legacy_1.001"),
    "odd_aniso": dict(shape=(15, 15, 17), L=(22.5, 22.5, 25.5), nb=7, log=True, flags=(True, True, False), rule="legacy_1.001"),
    "inscribed_rule": dict(shape=(24, 24, 18), L=(36.0, 36.0, 27.0), nb=10, log=False, flags=(True, False, True), rule="inscribed"),
    "prime_z": dict(shape=(12, 12, 37), L=(18.0, 18.0, 55.5), nb=6, log=True, flags=(True, False, False), rule="legacy_1.001"),
}


def make_box(case, seed=0):
    """float32 box, mean ~20, sigma ~10, from a 64-bit integer hash of the cell index."""
    n = int(np.prod(case["shape"]))
    off = np.uint64((int(seed) * 0x9E3779B97F4A7C15) % 2**64)
    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.uint64) + off
        x = i * np.uint64(0xD1342543DE82EF95) + np.uint64(0x2545F4914F6CDD1D)
        x ^= x >> np.uint64(32)
        x *= np.uint64(0xDA942042E4DD58B5)
        x ^= x >> np.uint64(29)
        y = x * np.uint64(0x9FB21C651E98DF25)
        y ^= y >> np.uint64(31)
    u1 = ((x >> np.uint64(11)).astype(np.float64) + 0.5) / 2.0**53
    u2 = ((y >> np.uint64(11)).astype(np.float64) + 0.5) / 2.0**53
    g = np.sqrt(-2.0 * np.log(u1)) * np.cos(2 * np.pi * u2)
    return (20.0 + 10.0 * g).astype(np.float32).reshape(case["shape"])


def main():
    from oracle import ps_oracle as O

    out = {}
    for name, case in CASES.items():
        box = make_box(case)
        p, k, sw = O.compute_power_lightcone(
            box, case["L"], case["nb"], log_bins=case["log"], ignore_kperp_zero=case["flags"][0],
            ignore_kpar_zero=case["flags"][1], ignore_k_zero=case["flags"][2], edge_rule=case["rule"],
            return_counts=True)
        out[f"{name}_power"] = p
        out[f"{name}_k"] = k
        out[f"{name}_counts"] = sw.astype(np.int64)
        out[f"{name}_boxsum"] = np.array([box.astype(np.float64).sum()])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz")
    np.savez(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
