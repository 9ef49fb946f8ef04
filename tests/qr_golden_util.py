"""Helpers for the QR golden vectors (tests/golden/lair_qr_golden.json)."""
import json
import os

import numpy as np

_DT = {"f32": np.float32, "f64": np.float64, "c64": np.complex64, "c128": np.complex128}
QR_GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lair_qr_golden.json")


def load_qr_golden():
    with open(QR_GOLDEN_PATH) as f:
        return json.load(f)


def mat(case, key):
    dt = _DT[case["dtype"]]
    if key in case:
        return np.array(case[key], dtype=dt)
    if key + "_re" in case:
        return (np.array(case[key + "_re"], dtype=np.float64) + 1j * np.array(case[key + "_im"], dtype=np.float64)).astype(dt)
    return None


def cvec(case, key):
    """[[re, im], ...] -> complex / real vector of the case's dtype."""
    dt = _DT[case["dtype"]]
    v = np.array([complex(p[0], p[1]) for p in case[key]], dtype=np.complex128)
    return v.astype(dt) if np.issubdtype(dt, np.complexfloating) else v.real.astype(dt)


def check_geqrf_case(case, qr, tau):
    eps = case["eps"]
    exp = mat(case, "qr")
    if exp is not None:
        assert np.max(np.abs(np.asarray(qr) - exp)) <= eps, (case["name"], qr)
    exp_tau = mat(case, "tau")
    if exp_tau is not None:
        assert tau.shape == exp_tau.shape, (case["name"], tau.shape)
        assert np.max(np.abs(tau - exp_tau)) <= eps, (case["name"], tau)


def check_qr_case(case, q, r):
    eps = case["eps"]
    eq, er = mat(case, "q"), mat(case, "r")
    assert q.shape == eq.shape and r.shape == er.shape, (case["name"], q.shape, r.shape)
    assert np.max(np.abs(q - eq)) <= eps, (case["name"], q)
    assert np.max(np.abs(r - er)) <= eps, (case["name"], r)


def qr_errors(a0, q, r):
    """(||A - Q R||_F / (max(m, n) eps ||A||_F), ||Q^H Q - I||_F / (m eps))."""
    m, n = a0.shape
    real = np.finfo(a0.dtype).dtype
    eps = np.finfo(real).eps / 2
    wide = np.complex128 if np.iscomplexobj(a0) else np.float64
    a, qq, rr = a0.astype(wide), q.astype(wide), r.astype(wide)
    na = np.linalg.norm(a)
    fact = float(np.linalg.norm(a - qq @ rr) / (max(m, n) * eps * na)) if na > 0 else 0.0
    orth = float(np.linalg.norm(qq.conj().T @ qq - np.eye(m)) / (m * eps))
    return fact, orth
