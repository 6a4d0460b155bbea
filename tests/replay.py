"""Replay helpers shared by the golden / parity tests: rebuild a recorded input on a given
engine, apply its view chain, make the recorded call through pdl_b200's operator surface."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

import pdl_b200 as P
from pdl_b200 import types as T, ufunc, ops, bad, basic
from pdl_b200.engine import PDLError

GOLDEN = Path(__file__).resolve().parent / "golden"
TYPE_ID = {n: i for i, n in enumerate(T.NAMES)}


def load_cases(fname):
    return json.loads((GOLDEN / fname).read_text())["cases"]


def build_input(spec, engine):
    if "scalar" in spec:
        v = spec["scalar"]
        return int(v) if spec.get("is_int") else float(v)
    t = TYPE_ID[spec["type"]]
    raw = np.frombuffer(bytes.fromhex(spec["hex"]), dtype=T.NP_DTYPE[t])
    p = P.PDL.from_numpy(raw.reshape(list(reversed(spec["dims"]))), t, engine)
    p.badflag = bool(spec["badflag"])
    bv = np.frombuffer(bytes.fromhex(spec["badvalue_hex"]), dtype=T.NP_DTYPE[t])[0]
    default = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
    if bv.tobytes() != default.tobytes():
        p.set_badvalue(complex(bv) if t in (T.CF, T.CD) else float(bv) if t in (T.F, T.D) else int(bv))
    for v in spec.get("views", []):
        p = getattr(p, v[0])(*v[1:])
    return p


def run_call(call, args):
    kind = call["kind"]
    if kind == "biop":
        a, b = args
        if call.get("inplace"):
            a = a.copy().inplace()
        return P.run_biop(call["op"], a, b, None, call.get("swap", 0))
    if kind == "ufunc":
        op = {"abs": "_rabs"}.get(call["op"], call["op"])
        return P.run_ufunc(op, args[0])
    if kind == "reduce":
        return getattr(ufunc, call["op"])(args[0])
    if kind == "whole":
        return getattr(ufunc, call["op"])(args[0])
    if kind == "matmult":
        return P.matmult(args[0], args[1])
    if kind == "ipow":
        return ops.ipow(args[0], args[1])
    if kind == "convert":
        return args[0].convert(TYPE_ID[call["to"]])
    if kind == "badop":
        f = getattr(bad, call["op"])
        if call.get("inplace"):
            a = args[0].copy()
            a.badflag = args[0].badflag
            f(a.inplace(), *args[1:])
            return a
        return f(*args)
    if kind == "axis":
        return getattr(basic, call["op"])(args[0])
    if kind == "sequence":
        return basic.sequence(TYPE_ID[call["type"]] if call.get("type") else None, *call["dims"], engine=call["_engine"])
    if kind == "inner":
        return P.inner(args[0], args[1])
    if kind == "outer":
        return P.outer(args[0], args[1])
    if kind == "minmaximum":
        return list(ufunc.minmaximum(args[0]))
    if kind == "n_ind":
        return getattr(ufunc, call["op"])(args[0], call["m"])
    raise ValueError(kind)


def ulp_diff(got: np.ndarray, want: np.ndarray) -> int:
    """Largest distance in units-in-the-last-place between two float arrays of equal dtype;
    identical NaN-ness and infinities are required."""
    assert got.dtype == want.dtype
    it = np.int32 if got.dtype == np.float32 else np.int64
    g, w = got.reshape(-1), want.reshape(-1)
    nan_g, nan_w = np.isnan(g), np.isnan(w)
    if not np.array_equal(nan_g, nan_w):
        return 1 << 62
    gi, wi = g.view(it).astype(np.int64), w.view(it).astype(np.int64)
    # map the sign-magnitude float ordering onto a monotone integer line
    mn = np.int64(np.iinfo(it).min)
    gi = np.where(gi < 0, mn - gi, gi)
    wi = np.where(wi < 0, mn - wi, wi)
    d = np.abs(gi - wi)
    d[nan_g] = 0
    return int(d.max()) if d.size else 0


def check_case(case, engine):
    """Replay one recorded case on `engine`; assert type, dims, badflag and values."""
    args = [build_input(s, engine) for s in case["inputs"]]
    case["call"]["_engine"] = engine
    if "error" in case:
        try:
            run_call(case["call"], args)
        except PDLError as e:
            want = case["error"]
            key = ("Mismatched implicit broadcast dimension" if "Mismatched" in want else
                   "index 'n' size 3, but ndarray dim has size 4" if "index 'n'" in want else
                   "m_size > n_size" if "m_size > n_size" in want else "Dim mismatch in matmult")
            assert key in str(e), (str(e), want)
            if "Dim mismatch" in want:
                assert str(e).strip() == want.strip()
            return
        raise AssertionError(f"{case['name']}: reference raised {case['error']!r}, we did not")
    out = run_call(case["call"], args)
    if "outputs" in case:
        assert len(out) == len(case["outputs"])
        for k, (o, want) in enumerate(zip(out, case["outputs"])):
            assert o.type == want["type"], (case["name"], k, o.type, want["type"])
            assert o.dims == want["dims"], (case["name"], k, o.dims, want["dims"])
            assert int(o.badflag) == want["badflag"], (case["name"], k, "badflag", o.badflag, want["badflag"])
            exp = np.frombuffer(bytes.fromhex(want["hex"]), dtype=T.NP_DTYPE[o.datatype])
            assert o.to_numpy().reshape(-1).tobytes() == exp.tobytes(), (case["name"], k, o.to_numpy(), exp)
        return
    want = case["output"]
    assert out.type == want["type"], (case["name"], out.type, want["type"])
    assert out.dims == want["dims"], (case["name"], out.dims, want["dims"])
    assert int(out.badflag) == want["badflag"], (case["name"], "badflag", out.badflag, want["badflag"])
    dt = T.NP_DTYPE[out.datatype]
    got = out.to_numpy().reshape(-1)
    exp = np.frombuffer(bytes.fromhex(want["hex"]), dtype=dt)
    tol = case.get("tol_ulp")
    # float sum/product/average: a NaN result is NaN on both sides, but its sign/payload is an
    # artefact of summation order and of x86-vs-GPU NaN generation, not of PDL semantics
    nan_free = (case["call"]["kind"] in ("reduce", "whole") and any(
        k in case["call"]["op"] for k in ("sum", "prod", "aver", "avg", "magn"))) or case["call"]["kind"] == "inner"
    if dt.kind == "c":
        # complex results: NaN parts must be NaN on both sides (payload / sign of a NaN is an x86-vs-GPU artefact),
        # every other part bit for bit
        ft = np.float32 if dt == np.complex64 else np.float64
        d = ulp_diff(got.view(ft), exp.view(ft))
        assert d == 0, (case["name"], f"{d} ulp", got, exp)
    elif dt.kind == "f" and (tol or nan_free):
        tol = tol or 0
        d = ulp_diff(got, exp)
        assert d <= tol, (case["name"], f"{d} ulp > {tol}", got, exp)
    else:
        assert got.tobytes() == exp.tobytes(), (case["name"], got, exp)
