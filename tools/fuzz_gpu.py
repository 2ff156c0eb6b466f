"""Randomised differential run: library (specialised kernels where available) vs the C oracle, many seeds.
usage: python tools/fuzz_gpu.py [seconds]   — prints one JSON line with the counts; exits 1 on the first mismatch.
Test infrastructure (uses oracle/)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import oracle
from qublas_b200 import abi, ops
from qublas_b200.qtypes import QType, TREE_VARIADIC, TREE_VECTOR, resolve_bcm, resolve_rc, resolve_tf


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def words(rng, t, n):
    return rng.integers(t.raw_min, t.raw_max + 1, size=n, dtype=np.int64).astype(oracle.np_dtype(t.word_bytes))


def rt(rng, lo=2, hi=8, flo=0, fhi=9, wide=False):
    return QType(int(rng.integers(lo, hi)) + (16 if wide else 0), int(rng.integers(flo, fhi)), bool(rng.random() < 0.9), int(rng.integers(0, 7)), int(rng.integers(0, 4)))


def levels(rng, elem, n):
    lv, i, f = [], elem.i, elem.f
    for _ in range(n):
        i += int(rng.integers(0, 3)); f -= int(rng.integers(0, 3))
        if i + f < 1:
            f = 1 - i
        lv.append(QType(i, f, bool(rng.random() < 0.9), int(rng.integers(0, 7)), int(rng.integers(0, 4))))
    return lv


def tdt(wb):
    return torch.int32 if wb == 4 else torch.int64


def one_gemul(rng, stats):
    M, N, K = (int(rng.choice(v)) for v in ([1, 2, 3, 4, 8, 12, 16, 20, 32, 33, 64], [1, 2, 4, 5, 8, 16, 32, 64], [1, 2, 7, 8, 16, 24, 31, 32, 48, 64, 100, 128]))
    wide = rng.random() < 0.2
    a, b = rt(rng, 3, 8, 2, 9, wide), rt(rng, 3, 8, 2, 9, wide)
    mul = QType(a.i + b.i - int(rng.integers(0, 3)), max(a.f, b.f) - int(rng.integers(0, 3)), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
    lv = levels(rng, mul, int(rng.choice([0, 1, 3, 7])))
    last = lv[-1] if lv else mul
    c = QType(last.i + int(rng.integers(-1, 3)), max(last.f - int(rng.integers(0, 3)), 1 - last.i), True, int(rng.integers(0, 7)), int(rng.integers(0, 4)))
    form = TREE_VECTOR if rng.random() < 0.5 else TREE_VARIADIC
    plan = ops.make_gemul_plan(a, b, c, M, N, K, mul=mul, add=lv or None, form=form, trans_a=bool(rng.random() < 0.3), trans_b=bool(rng.random() < 0.3))
    batch = int(rng.integers(1, 6))
    A, B = words(rng, a, batch * M * K), words(rng, b, batch * K * N)
    Cd = torch.empty(batch * M * N, dtype=tdt(c.word_bytes), device="cuda")
    ops.Qgemul(Cd, dev(A), dev(B), plan, batch=batch)
    want = oracle.gemul_batched(plan, A, B, batch)
    stats[ops.plan_info_gemul(plan).kernel.decode()] = stats.get(ops.plan_info_gemul(plan).kernel.decode(), 0) + 1
    return np.array_equal(Cd.cpu().numpy(), want), ("gemul", M, N, K, form, str(a), str(b), str(mul), [str(x) for x in lv], str(c))


def one_cgemul(rng, stats):
    M, N, K = (int(rng.choice(v)) for v in ([1, 2, 3, 5, 8, 16, 32, 64], [1, 2, 4, 7], [1, 2, 5, 8, 16, 24, 33, 64, 128]))
    kind = str(rng.choice(["tf", "basic", "real_a", "real_b"]))
    tile_shape = rng.random() < 0.25                      # a quarter of the cases: shapes the register-tile kernels take
    if tile_shape:
        M, N, K = (int(rng.choice(v)) for v in ([32, 64, 128], [16, 32, 64], [16, 32, 64]))
        kind = str(rng.choice(["tf", "basic"]))
    ar, ai, br, bi = rt(rng, 2, 7, 1, 9), rt(rng, 2, 7, 1, 9), rt(rng, 2, 7, 1, 9), rt(rng, 2, 7, 1, 9)
    def tg(names):
        return {k: rt(rng, 5, 12, 2, 10) for k in names if rng.random() < 0.8}
    if kind == "tf":
        tags = tg(("abT", "cdT", "abcT", "cdbT", "badT", "ABT", "BCT")); r = resolve_tf(ar, ai, br, bi, tags); leaf = (r["AB"], r["BC"]); args = (ar, ai, br, bi)
    elif kind == "basic":
        tags = tg(("acT", "bdT", "adT", "bcT", "acbdT", "adbcT")); r = resolve_bcm(ar, ai, br, bi, tags); leaf = (r["acbd"], r["adbc"]); args = (ar, ai, br, bi)
    elif kind == "real_a":
        tags = {"realT": rt(rng, 6, 12, 2, 10), "imagT": rt(rng, 6, 12, 2, 10)}; r = resolve_rc(True, ar, br, bi, tags); leaf = (r["ac"], r["ad"]); args = (ar, None, br, bi)
    else:
        tags = {"realT": rt(rng, 6, 12, 2, 10), "imagT": rt(rng, 6, 12, 2, 10)}; r = resolve_rc(False, br, ar, ai, tags); leaf = (r["ac"], r["bc"]); args = (ar, ai, br, None)
    lre, lim = levels(rng, leaf[0], int(rng.choice([1, 3, 6]))), levels(rng, leaf[1], int(rng.choice([1, 2, 5])))
    form = TREE_VECTOR if rng.random() < 0.5 else TREE_VARIADIC
    plan = ops.make_cgemul_plan(*args, rt(rng, 8, 12, 2, 6), rt(rng, 8, 12, 2, 6), M, N, K, tf_tags=tags, add_re=lre, add_im=lim, form=form, product=kind,
                                trans_a=bool(rng.random() < 0.3) and not tile_shape, trans_b=bool(rng.random() < 0.3) and not tile_shape)
    batch = int(rng.integers(1, 5))
    def cw(re_t, im_t, n):
        if im_t is None:
            return words(rng, re_t, n).astype(np.int32)
        return np.stack([words(rng, re_t, n), words(rng, im_t, n)], axis=1).astype(np.int32).reshape(-1)
    A, B = cw(args[0], args[1], batch * M * K), cw(args[2], args[3], batch * K * N)
    Cd = torch.empty(batch * M * N * 2, dtype=torch.int32, device="cuda")
    ops.Qcgemul(Cd, dev(A), dev(B), plan, batch=batch)
    want = oracle.cgemul_batched(plan, A, B, batch)
    k = ops.plan_info_cgemul(plan).kernel.decode(); stats[k] = stats.get(k, 0) + 1
    return np.array_equal(Cd.cpu().numpy(), want), ("cgemul", kind, M, N, K, form)


def one_reduce(rng, stats):
    n = int(rng.choice([1, 2, 3, 63, 64, 65, 100, 255, 256, 257, 511, 1000, 4095, 4096, 4097, 9999, 65536, 65537, 70001, 1 << 17, (1 << 18) + 3]))
    wide = rng.random() < 0.2
    elem = rt(rng, 3, 10, 2, 9, wide)
    lv = levels(rng, elem, int(rng.choice([0, 1, 3, 8, 20])))
    form = TREE_VECTOR if (rng.random() < 0.8 or n > 5000) else TREE_VARIADIC
    batch = int(rng.integers(1, 4))
    x = words(rng, elem, batch * n)
    plan = ops.make_reduce_plan(elem, lv or None, n, form=form)
    out, _ = ops.Qreduce(dev(x), elem, lv or None, length=n, batch=batch, form=form)
    want, _ = oracle.reduce_batched(plan, x, batch)
    k = ops.plan_info_reduce(plan).kernel.decode(); stats[k] = stats.get(k, 0) + 1
    return np.array_equal(out.cpu().numpy(), want), ("reduce", n, form, str(elem), [str(x) for x in lv])


def one_ew(rng, stats):
    n = int(rng.choice([1, 5, 1000, 4099, 100003]))
    wide = rng.random() < 0.3
    a, b = rt(rng, 3, 9, 0, 9, wide), rt(rng, 3, 9, 0, 9, wide)
    op = str(rng.choice(["mul", "add", "sub"]))
    to = rt(rng, 4, 14, 0, 10, wide)
    x, y = words(rng, a, n), words(rng, b, n)
    out, _ = ops.elementwise(op, dev(x), dev(y), a, b, to=to)
    want = oracle.elementwise(ops.make_ew_plan(op, a, b, n, to), x, y)
    stats["ew"] = stats.get("ew", 0) + 1
    return np.array_equal(out.cpu().numpy(), want), ("ew", op, n, str(a), str(b), str(to))


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "12345")))
    t0 = time.time(); done = refused = 0; stats = {}
    fns = [one_gemul, one_cgemul, one_reduce, one_ew]
    while time.time() - t0 < budget:
        fn = fns[done % len(fns)]
        try:
            ok, what = fn(rng, stats)
        except abi.QbError as e:
            if e.code != abi.QB_EDOMAIN:
                raise
            refused += 1; done += 1
            continue
        if not ok:
            print(json.dumps({"mismatch": what})); sys.exit(1)
        done += 1
    print(json.dumps({"cases": done, "refused_edomain": refused, "kernels": stats, "seconds": round(time.time() - t0, 1)}))


if __name__ == "__main__":
    main()
