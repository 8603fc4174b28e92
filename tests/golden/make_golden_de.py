"""Goldens for the DE and CMA searches (src/opt/de.cpp, src/opt/cma.cpp) recorded from the reference's own OptDE / OptCMA through
oracle/_ref/libsacref_nc.so (ref_harness.cpp: ref_de_run / ref_cma_run, one evaluation thread, so the traces are in the
reference's evaluation order).
usage: python tests/golden/make_golden_de.py   (needs /root/reference via oracle/_ref; writes tests/golden/golden_de.json)"""
import hashlib, json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import oracle_lib as ol


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def cost_of(x, xmin, xmax, ties):
    z = (x - xmin) / (xmax - xmin)
    c = float(np.sum((z - 0.37) ** 2) + 0.05 * np.sum(np.cos(9 * z)))
    return float(np.floor(c * 8)) if ties else c      # integer-valued costs produce ties, as the bitplane objective does


def main():
    ref = ol.ref_lib(nc=True)
    vmin, vmax, vdef = ol.base_profile()
    idx = [i for i in range(58) if i not in (56, 57)]
    xmin = vmin[idx].astype(np.float64); xmax = vmax[idx].astype(np.float64); xs = vdef[idx].astype(np.float64)
    out = {"generator": "tests/golden/make_golden_de.py", "cases": []}
    for nfunc, sigma, ties in ((45, 0.15, False), (200, 0.25, False), (331, 0.2, True), (17, 0.2, False)):
        trace = []

        def cb(xp, n, _u):
            x = np.ctypeslib.as_array(xp, shape=(n,)).copy()
            trace.append(x)
            return cost_of(x, xmin, xmax, ties)

        fn = ol.COST_CB(cb)
        xb = np.zeros(56)
        fb = ref.ref_de_run(56, ol._p(xmin, ol._f64p), ol._p(xmax, ol._f64p), ol._p(xs, ol._f64p), nfunc, sigma, fn, None, ol._p(xb, ol._f64p))
        out["cases"].append(dict(nfunc=nfunc, sigma=sigma, ties=ties, best=float(fb), xbest_sha1=sha(xb), trace_sha1=sha(np.stack(trace)),
                                 evals=len(trace)))
    out["cma"] = []
    for nfunc, sigma, ties in ((80, 0.0, False), (400, 0.1, False), (250, 0.0, True)):
        trace = []

        def cb2(xp, n, _u):
            x = np.ctypeslib.as_array(xp, shape=(n,)).copy()
            trace.append(x)
            return cost_of(x, xmin, xmax, ties)

        fn = ol.COST_CB(cb2)
        xb = np.zeros(56)
        fb = ref.ref_cma_run(56, ol._p(xmin, ol._f64p), ol._p(xmax, ol._f64p), ol._p(xs, ol._f64p), nfunc, sigma, fn, None, ol._p(xb, ol._f64p))
        out["cma"].append(dict(nfunc=nfunc, sigma=sigma, ties=ties, best=float(fb), xbest_sha1=sha(xb), trace_sha1=sha(np.stack(trace)),
                               evals=len(trace)))
    json.dump(out, open(os.path.join(HERE, "golden_de.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
