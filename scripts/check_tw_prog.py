#!/usr/bin/env python3
"""Big-integer model of jac_mul_prog (go-eth-kzg_b200/csrc/fk20.cuh): the odd-multiples table is
built on a curve isomorphic to E (scaled so that 2P is affine), brought to one common Z, used
as an affine table, and the result is mapped back by multiplying Z.  Checks the algebra against
plain double-and-add for a few twiddles.  Development aid; not used at run time."""
import importlib.util, os, sys, io, contextlib
here = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("gc", os.path.join(here, "gen_constants.py"))
gc = importlib.util.module_from_spec(spec)
with contextlib.redirect_stdout(io.StringIO()):
    spec.loader.exec_module(gc)
p, r, beta = gc.p, gc.r, gc.beta


def jdbl(P):
    X, Y, Z = P
    A = X * X % p; B = Y * Y % p; C = B * B % p
    D = 2 * ((X + B) ** 2 - A - C) % p
    E = 3 * A % p; F = E * E % p
    X3 = (F - 2 * D) % p
    return (X3, (E * (D - X3) - 8 * C) % p, 2 * Y * Z % p)


def jmadd(P, q, raw=False):
    """Jacobian + affine, returns (sum, zr) with Z3 = Z1 * zr; raw = no infinity case (table build)"""
    X1, Y1, Z1 = P
    if Z1 == 0 and not raw: return (q[0], q[1], 1), None
    ZZ = Z1 * Z1 % p
    U2 = q[0] * ZZ % p; S2 = q[1] * Z1 * ZZ % p
    H = (U2 - X1) % p; rr = (S2 - Y1) % p
    HH = H * H % p; HHH = H * HH % p; V = X1 * HH % p
    X3 = (rr * rr - HHH - 2 * V) % p
    return (X3, (rr * (V - X3) - Y1 * HHH) % p, Z1 * H % p), H


def to_aff(P):
    X, Y, Z = P
    if Z == 0: return None
    zi = pow(Z, -1, p)
    return (X * zi * zi % p, Y * zi ** 3 % p)


def jac_mul_prog(P, ops, trailing):
    X, Y, Z = P
    XD, YD, C = jdbl(P)
    C2 = C * C % p; C3 = C2 * C % p
    T = [(X * C2 % p, Y * C3 % p, Z)]
    zr = [None]
    for k in range(1, 8):
        t, h = jmadd(T[-1], (XD, YD), raw=True)
        T.append(t); zr.append(h)
    ZL = T[7][2]
    tab = [None] * 8
    s = 1
    for k in range(7, -1, -1):
        s2 = s * s % p
        tab[k] = (T[k][0] * s2 % p, T[k][1] * s2 * s % p)
        if k: s = s * zr[k] % p
    acc = (0, 0, 0)
    for op in ops:
        for _ in range(op & 255): acc = jdbl(acc)
        e = tab[(op >> 8) & 7]
        if (op >> 12) & 1: e = (beta * e[0] % p, e[1])
        if (op >> 11) & 1: e = (e[0], (-e[1]) % p)
        acc, _ = jmadd(acc, e)
    for _ in range(trailing): acc = jdbl(acc)
    return (acc[0], acc[1], acc[2] * ZL % p * C % p)


Q = gc.ec_mul(0xdeadbeefcafe, gc.G)
z = 0x123456789abcdef
Pj = (Q[0] * z * z % p, Q[1] * z ** 3 % p, z)
for j in (1, 3, 32, 63, 65, 100, 127):
    ops, tr = gc.progs[j]
    got = to_aff(jac_mul_prog(Pj, ops, tr))
    assert got == gc.ec_mul(pow(gc.w128, j, r), Q), j
assert jac_mul_prog((0, 0, 0), *gc.progs[5])[2] == 0   # infinity in -> Z = 0 out (C = Z_L = 0)
print("jac_mul_prog model OK")
