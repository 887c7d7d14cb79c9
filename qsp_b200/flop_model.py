"""Algorithmic flop / byte model of the reference's sweep sequence:
the contract figure F_ref the roofline fraction is quoted on.  complex128, one complex
multiply-add = 8 flops; textbook counts QR(m,n)=8(2MN^2-2/3 N^3), GEMM(m,k,n)=8mkn,
SVD(m,n)=4(6MN^2+20N^3) with M=max(m,n), N=min(m,n)."""


def _qr(m, n):
    M, N = max(m, n), min(m, n)
    return 8.0 * (2.0 * M * N * N - (2.0 / 3.0) * N ** 3)


def _gemm(m, k, n):
    return 8.0 * m * k * n


def _svd(m, n):
    M, N = max(m, n), min(m, n)
    return 4.0 * (6.0 * M * N * N + 20.0 * N ** 3)


def bond_profile(L, chi):
    return [min(chi, 2 ** min(i, 62), 2 ** min(L - i, 62)) for i in range(L + 1)]


def grown_profile(L, beta):
    """transient bonds after the inverse layer (right-to-left gate_split): b'_i = min(2 b_{i-1}, 2 b'_{i+1}, 2^i, 2^(L-i))"""
    bp = list(beta)
    bp[L] = 1
    for i in range(L - 1, 0, -1):
        bp[i] = min(2 * beta[i - 1], 2 * bp[i + 1], 2 ** min(i, 62), 2 ** min(L - i, 62))
    return bp


def _left_canon(L, b):
    f = 0.0
    for i in range(L - 1):
        f += _qr(2 * b[i], b[i + 1]) + _gemm(b[i + 1], b[i + 1], 2 * b[i + 2])
    return f


def _right_canon(L, b):
    f = 0.0
    for i in range(L - 1, 0, -1):
        f += _qr(2 * b[i + 1], b[i]) + _gemm(2 * b[i - 1], b[i], b[i])
    return f


def _right_compress(L, b, max_bond):
    """quimb tensor_compress_bond per bond: QR(T1), LQ(T2), M = R L, SVD(M), 2 GEMMs."""
    f = 0.0
    r = b[L]
    for i in range(L - 1, 0, -1):
        l1, bb = b[i - 1], b[i]
        k1 = min(2 * l1, bb)
        k2 = min(bb, 2 * r)
        f += _qr(2 * l1, bb) + _qr(2 * r, bb) + _gemm(k1, bb, k2) + _svd(k1, k2)
        k = min(k1, k2, max_bond if max_bond else 10 ** 9)
        f += _gemm(2 * l1, k1, k) + _gemm(k, k2, 2 * r)
        r = k
    return f


def flops_per_mps(L, chi, layers, compression=True, bonds=None):
    """F_ref = layers * (S1 + S2 + S3 + S4) for an MPS with bond profile `bonds` (default: the random-MPS profile
    min(chi, 2^i, 2^(L-i)); pass the live profile of a structured state so that the model is the work)."""
    b = [min(int(x), chi) for x in bonds] if bonds is not None else bond_profile(L, chi)
    bp = grown_profile(L, b)
    s1 = _left_canon(L, b) + _right_compress(L, b, 2)
    s2 = 0.0
    for i in range(L - 1):
        s2 += _gemm(2 * b[i], b[i + 1], 2 * bp[i + 2]) + _svd(2 * b[i], 2 * bp[i + 2])
    s3 = (_right_canon(L, bp) + _left_canon(L, bp) + _right_compress(L, bp, chi)) if compression else 0.0
    bb = b if compression else bp
    s4 = _right_canon(L, bb) + _left_canon(L, bb) + _right_compress(L, bb, None)
    return {"S1": s1, "S2": s2, "S3": s3, "S4": s4, "per_layer": s1 + s2 + s3 + s4,
            "total": layers * (s1 + s2 + s3 + s4)}


def bytes_per_mps(L, chi, layers, bonds=None):
    """B_ref: every sweep reads and writes every site tensor once."""
    b = [min(int(x), chi) for x in bonds] if bonds is not None else bond_profile(L, chi)
    bp = grown_profile(L, b)
    T = sum(32.0 * b[i] * b[i + 1] for i in range(L))
    Tp = sum(32.0 * bp[i] * bp[i + 1] for i in range(L))
    return layers * (5 * T + (T + Tp) + (5 * Tp + T) + 6 * T)
