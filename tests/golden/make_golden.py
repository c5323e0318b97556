#!/usr/bin/env python
"""Generates tests/golden/*.json.

The reference (Julia 0.3) cannot be executed here, so the fixtures are of two kinds:
  * `reference_kats.json` -- numbers COPIED from the reference's own tests / example output (file:line cited
    per entry) and the hand-derived enumerations of SURVEY.md section 4.  They pin the oracle.
  * `oracle_vectors.json` -- seeded input/output vectors produced by the oracle AFTER it passed the KATs
    (ancestors of the literal src/rand.jl algorithm, grid resampling, one conditional-SMC run, one MH run).
    They freeze the oracle's behaviour so that a later change to oracle/ or to the CUDA path that alters
    any integer is caught even when both sides drift together.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from oracle import oracle as orc
    from gpu_util import cfg2_tables, cfg5_bnet, dyadic_weights
    orc.build()

    kats = {
        "enum_three_flips": {"cite": "test/inferencetests.jl:3-15, examples/variational.jl:22-26",
                             "p": {"0": 0.054615886286517, "1": 0.351538628762172, "2": 0.445384113713482, "3": 0.148461371237827}},
        "mem_flip": {"cite": "test/memtests.jl:7-11", "p_true": 0.7},
        "hellinger": {"cite": "test/erptests.jl:12-17", "p": [0.25, 0.25, 0.5], "q_counts": {"1": 2, "3": 3}, "value": 0.36885, "tol": 1e-5},
        "rand_nan": {"cite": "test/randtests.jl:1-2", "inputs": [["nan"], ["nan", "nan"]], "error": "unexpected nan in rand"},
        "zero_mass": {"cite": "test/inferencetests.jl:17-20, src/erp.jl:48", "error": "Distribution has zero probability mass."},
        "hmm_example_enum": {"cite": "examples/hmm.jl:4-32 enumerated by hand (SURVEY.md section 4)", "Z": 0.12916,
                             "p": {"FFF": 0.8296918550634872, "TFF": 0.09218798389594302, "FFT": 0.039509135955404145,
                                   "FTF": 0.01693248683803035, "TTF": 0.010243109321771441, "FTT": 0.004389903995044904,
                                   "TFT": 0.004389903995044904, "TTT": 0.0026556209352740765}},
        "hmm2_example_enum": {"cite": "examples/hmm2.jl:5-24 enumerated by hand (SURVEY.md section 4)",
                              "Z": 0.13253212222378208, "logZ": -2.0209302310601185,
                              "p": {"FTTTT": 0.5500553282992594, "FFTTT": 0.20235404678810578, "FFFFF": 0.09067153285758026,
                                    "FFFTT": 0.07444189365118824, "FFFFT": 0.027385642236143076, "FTTTF": 0.02248378297645619}},
    }
    json.dump(kats, open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1, sort_keys=True)

    rng = np.random.default_rng(20261019)
    vec = {}
    # literal src/rand.jl + src/pmcmc.jl:48-52 on seeded weights / uniforms
    w = rng.random(257)
    r = rng.random(300)
    vec["literal"] = {"w": w.tolist(), "r": r.tolist(), "anc": orc.resample_literal(w, r).tolist()}
    wd = dyadic_weights(rng, 1000)
    rd = rng.random(500)
    vec["literal_dyadic"] = {"w": wd.tolist(), "r": rd.tolist(), "anc": orc.resample_literal(wd, rd).tolist()}
    # grid resampling (extension) on log-weights
    lw = (rng.normal(size=5000) * 3.0)
    vec["systematic_log_f64"] = {"lw": lw.tolist(), "u": 0.3125, "m": 5000,
                                 "anc": orc.resample_systematic(lw, 0.3125, 5000, kind=1).tolist()}
    vec["multinomial_log_f64"] = {"lw": lw[:700].tolist(), "r": r.tolist(),
                                  "anc": orc.resample_multinomial(lw[:700], r, kind=1).tolist()}
    # conditional SMC (pmcmc) on the cfg2 model: marginal counts of 3 iterations x 50 particles, literal numerics
    A, logF, pi = cfg2_tables()
    m = orc.Hmm(A, logF, pi=pi)
    for mode, name in [(orc.LITERAL, "literal"), (orc.SYSTEMATIC, "systematic")]:
        out = orc.hmm_pmcmc(m, orc.opts(mode, seed=5), 3, 50, want_joint=False)
        vec["pmcmc_cfg2_" + name] = {"iters": 3, "N": 50, "seed": 5, "marg": out["marg"].tolist(), "logZ_first": out["logZ_first"]}
    # batched MH on the cfg5 Bayes net
    sites, factors, query = cfg5_bnet()
    b = orc.Bnet([s[0] for s in sites], [s[1] for s in sites], [f[0] for f in factors], [f[1] for f in factors],
                 sum(1 << q for q in query))
    counts, acc = orc.bnet_mh(b, 3, 64, 200)
    vec["mh_cfg5"] = {"seed": 3, "chains": 64, "steps": 200, "counts": counts.tolist(), "accepts": int(acc)}
    json.dump(vec, open(os.path.join(HERE, "oracle_vectors.json"), "w"))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
