""" development tool: drives tools/tc_probe2.cu with shared-memory layout / descriptor hypotheses for
 (C) an MN-major 128B-swizzled A operand and (D) an MN-major un-swizzled ("interleave") B operand, the two
 operand forms the tensor-core SlotAttention kernel needs.  Run on a B200: python tools/tc_probe2.py """
import ctypes, os, subprocess, sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libtcprobe2.so")
if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "tc_probe2.cu")):
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-shared", "-Xcompiler", "-fPIC", "-o", SO, os.path.join(HERE, "tc_probe2.cu"), "-lcuda"], check=True)
if "--build-only" in sys.argv:
    sys.exit(0)
CASE = None
for a in sys.argv[1:]:
    if a.startswith("--case="):
        CASE = a.split("=")[1]
if CASE is None:      # one subprocess per case: a faulting hypothesis must not take the others down
    for cs in ("A", "C", "D", "E"):
        subprocess.run([sys.executable, os.path.abspath(__file__), "--case=" + cs])
    sys.exit(0)
lib = ctypes.CDLL(SO)
lib.probe_mma.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                          ctypes.c_int, ctypes.c_uint, ctypes.c_void_p, ctypes.c_int]
lib.probe_mma.restype = ctypes.c_int

REGION = 64 * 1024
rng = np.random.default_rng(0)


def desc(addr, lbo, sbo, layout):
    return ((addr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | (layout << 61)


def idesc(M, N, a_mn=0, b_mn=0):
    return (1 << 4) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def sw128(row, byte):
    chunk = byte >> 4
    return row * 128 + (((chunk ^ (row & 7)) << 4) | (byte & 15))


def run(imgA, imgB, ad, bd, idsc, ncols):
    a = torch.from_numpy(imgA.view(np.uint8)).cuda()
    b = torch.from_numpy(imgB.view(np.uint8)).cuda()
    adt = torch.tensor(np.array(ad, dtype=np.uint64).view(np.int64), device="cuda")
    bdt = torch.tensor(np.array(bd, dtype=np.uint64).view(np.int64), device="cuda")
    out = torch.zeros(128, ncols, device="cuda")
    r = lib.probe_mma(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), adt.data_ptr(), bdt.data_ptr(), len(ad),
                      idsc, out.data_ptr(), ncols)
    if r != 0:
        print("kernel failed", r)
        sys.exit(2)
    return out.cpu().numpy()


def put(img, off, val):
    img[off >> 1] = val


def report(name, got, ref):
    err = np.abs(got - ref).max()
    ok = err < 2e-2
    print(f"{name:70s} max|err| {err:9.3e} (max|ref| {np.abs(ref).max():.2f}) {'OK' if ok else 'WRONG'}")
    return ok


def kmajor_sw128_image(mat, block_rows):
    """ mat [rows][K] fp16 -> image of [K/64 blocks][rows][128 B] swizzled; block stride = block_rows*128 """
    rows, K = mat.shape
    img = np.zeros(REGION // 2, dtype=np.float16)
    for r in range(rows):
        for k in range(K):
            blk, kk = divmod(k, 64)
            put(img, blk * block_rows * 128 + sw128(r, kk * 2), mat[r, k])
    return img


results = {}
# ---- case A/B: proven K-major path with small N
for N in ((16, 8) if CASE == "A" else ()):
    A = rng.standard_normal((128, 64)).astype(np.float16)
    Bm = rng.standard_normal((N, 64)).astype(np.float16)
    ia, ib = kmajor_sw128_image(A, 128), kmajor_sw128_image(Bm, N)
    ad = [desc(k * 32, 16, 1024, 2) for k in range(4)]
    bd = [desc(k * 32, 16, 1024, 2) for k in range(4)]
    got = run(ia, ib, ad, bd, idesc(128, N), N)
    report(f"A: K-major SW128 both, N={N}", got, A.astype(np.float32) @ Bm.astype(np.float32).T)

# ---- case C: A MN-major SW128.  c[t][m]: memory [half h][token t][64 features], rows of 128 B swizzled by t
T = 128
c = rng.standard_normal((T, 128)).astype(np.float16)           # [token][feature]
N = 16
Bw = rng.standard_normal((N, T)).astype(np.float16)            # [n][token]  (K-major, K = tokens)
imgA = np.zeros(REGION // 2, dtype=np.float16)
for t in range(T):
    for m in range(128):
        h, mm = divmod(m, 64)
        put(imgA, h * T * 128 + sw128(t, mm * 2), c[t, m])
imgB = kmajor_sw128_image(Bw, N)
ref = c.astype(np.float32).T @ Bw.astype(np.float32).T          # [feature][n]
bd = [desc((j // 4) * N * 128 + (j % 4) * 32, 16, 1024, 2) for j in range(8)]
for name, lbo, sbo in ((("LBO=half stride, SBO=1024", T * 128, 1024), ("LBO=1024, SBO=half stride", 1024, T * 128)) if CASE == "C" else ()):
    ad = [desc(j * 16 * 128, lbo, sbo, 2) for j in range(8)]
    got = run(imgA, imgB, ad, bd, idesc(128, N, a_mn=1), N)
    results[("C", name)] = report(f"C: A MN-major SW128 ({name})", got, ref)

# ---- case D: B MN-major no-swizzle.  w[t][n]: memory [8-token group g][n-block nb][t%8][8 n] (128 B blocks)
for N in ((16, 8) if CASE == "D" else ()):
    NB = N // 8
    A2 = rng.standard_normal((128, T)).astype(np.float16)       # [m][token], K-major with K = tokens
    w = rng.standard_normal((T, N)).astype(np.float16)          # [token][n]
    imgA2 = kmajor_sw128_image(A2, 128)
    imgB2 = np.zeros(REGION // 2, dtype=np.float16)
    for t in range(T):
        for n in range(N):
            g, tt = divmod(t, 8)
            nb, nn = divmod(n, 8)
            put(imgB2, g * NB * 128 + nb * 128 + tt * 16 + nn * 2, w[t, n])
    ref = A2.astype(np.float32) @ w.astype(np.float32)
    ad = [desc((j // 4) * 128 * 128 + (j % 4) * 32, 16, 1024, 2) for j in range(8)]
    for name, lbo, sbo in (("LBO=k-group stride, SBO=n-block stride", NB * 128, 128), ("LBO=n-block stride, SBO=k-group stride", 128, NB * 128)):
        bd = [desc(j * 2 * NB * 128, lbo, sbo, 0) for j in range(8)]
        got = run(imgA2, imgB2, ad, bd, idesc(128, N, b_mn=1), N)
        results[("D", N, name)] = report(f"D: B MN-major no-swizzle N={N} ({name})", got, ref)

if CASE != "E":
    sys.exit(0)
# ---- case E: both together (MMA2 of the SlotAttention kernel), every LBO/SBO combination
N = 16
NB = 2
w = rng.standard_normal((T, N)).astype(np.float16)
imgB2 = np.zeros(REGION // 2, dtype=np.float16)
for t in range(T):
    for n in range(N):
        g, tt = divmod(t, 8)
        nb, nn = divmod(n, 8)
        put(imgB2, g * NB * 128 + nb * 128 + tt * 16 + nn * 2, w[t, n])
ref = c.astype(np.float32).T @ w.astype(np.float32)
for an, albo, asbo in (("A:LBO=half,SBO=1024", T * 128, 1024), ("A:LBO=1024,SBO=half", 1024, T * 128)):
    for bn, blbo, bsbo in (("B:LBO=kgrp,SBO=nblk", NB * 128, 128), ("B:LBO=nblk,SBO=kgrp", 128, NB * 128)):
        ad = [desc(j * 16 * 128, albo, asbo, 2) for j in range(8)]
        bd = [desc(j * 2 * NB * 128, blbo, bsbo, 0) for j in range(8)]
        got = run(imgA, imgB2, ad, bd, idesc(128, N, a_mn=1, b_mn=1), N)
        report(f"E: {an} {bn}", got, ref)
# N = 8 sub-operand of the same B image (hi block only): n-block stride irrelevant
ref8 = c.astype(np.float32).T @ w[:, :8].astype(np.float32)
for an, albo, asbo in (("A:LBO=half,SBO=1024", T * 128, 1024), ("A:LBO=1024,SBO=half", 1024, T * 128)):
    for bn, blbo, bsbo in (("B:LBO=kgrp,SBO=nblk", NB * 128, 128), ("B:LBO=nblk,SBO=kgrp", 128, NB * 128)):
        ad = [desc(j * 16 * 128, albo, asbo, 2) for j in range(8)]
        bd = [desc(j * 2 * NB * 128, blbo, bsbo, 0) for j in range(8)]
        got = run(imgA, imgB2, ad, bd, idesc(128, 8, a_mn=1, b_mn=1), 8)
        report(f"E8: {an} {bn}", got, ref8)
