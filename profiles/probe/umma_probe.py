import ctypes, os, sys
import numpy as np
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "umma_probe.so"))
lib.umma_probe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                           ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
NF = 8192  # floats per image (32 KB)

def idesc(a_mn, b_mn, negb=0):
    return (1 << 4) | (2 << 7) | (2 << 10) | (negb << 14) | (a_mn << 15) | (b_mn << 16) | ((128 >> 3) << 17) | ((128 >> 4) << 24)

def desc_hi(lbo_field_bytes, sbo_field_bytes, swz=0):
    return ((lbo_field_bytes >> 4) << 16) | ((sbo_field_bytes >> 4) << 32) | (1 << 46) | (swz << 61)

def run(Aimg, Bimg, dha, dhb, ide, nmma=1, astep=0, bstep=0):
    D = np.zeros((128, 128), np.float32)
    Aimg = np.ascontiguousarray(Aimg, np.float32); Bimg = np.ascontiguousarray(Bimg, np.float32)
    rc = lib.umma_probe(Aimg.ctypes.data, Bimg.ctypes.data, NF, dha, dhb, ide, nmma, astep, bstep, D.ctypes.data)
    assert rc == 0, rc
    return D

rng = np.random.default_rng(0)
A = rng.integers(-3, 4, (128, 8)).astype(np.float32)
B = rng.integers(-3, 4, (128, 8)).astype(np.float32)
ref = A @ B.T

def img_mn(X, SBO, LBO):
    im = np.zeros(NF, np.float32)
    for m in range(128):
        for k in range(X.shape[1]):
            off = (m // 4) * SBO + (k % 8) * 16 + (m % 4) * 4 + (k // 8) * LBO
            im[off // 4] = X[m, k]
    return im

def img_k(X, SBO, LBO):
    im = np.zeros(NF, np.float32)
    for m in range(128):
        for k in range(X.shape[1]):
            off = (m // 8) * SBO + (m % 8) * 16 + (k // 4) * LBO + (k % 4) * 4
            im[off // 4] = X[m, k]
    return im

for SBO, LBO in ((128, 4096), (144, 4608)):
    for name, dh in (("fields(LBO=K-stride,SBO=MN-stride)", desc_hi(LBO, SBO)), ("fields swapped", desc_hi(SBO, LBO))):
        D = run(img_mn(A, SBO, LBO), img_mn(B, SBO, LBO), dh, dh, idesc(1, 1))
        print("MN-major SBO=%d LBO=%d %s: max err %.1f  D[0,:4]=%s ref=%s" % (SBO, LBO, name, np.abs(D - ref).max(), D[0, :4], ref[0, :4]))
for SBO, LBO in ((256, 128), (128, 4096), (144, 4608)):
    for name, dh in (("fields(LBO=K-stride,SBO=MN-stride)", desc_hi(LBO, SBO)), ("fields swapped", desc_hi(SBO, LBO))):
        D = run(img_k(A, SBO, LBO), img_k(B, SBO, LBO), dh, dh, idesc(0, 0))
        print("K-major SBO=%d LBO=%d %s: max err %.1f  D[0,:4]=%s ref=%s" % (SBO, LBO, name, np.abs(D - ref).max(), D[0, :4], ref[0, :4]))

# scans: which (m) lights for a single 1.0 at byte offset X of A with B == 1 everywhere (MN-major idesc, my descriptor)
ones = np.ones(NF, np.float32)
dh = desc_hi(4096, 128)
for X in (0, 4, 8, 12, 16, 32, 112, 128, 256, 512, 1024, 2048, 4096):
    im = np.zeros(NF, np.float32); im[X // 4] = 1.0
    D = run(im, ones, dh, dh, idesc(1, 1))
    rows = np.nonzero(np.abs(D).sum(1))[0]
    print("A single @%d -> rows %s val %s" % (X, rows[:8], D[rows[0], 0] if len(rows) else None))
D = run(ones, ones, dh, dh, idesc(1, 1))
print("all ones: D unique", np.unique(D)[:5])
# K association: A single at X (row m known), B single at Y
for X, Y in ((0, 0), (0, 16), (16, 16), (0, 4), (16, 20)):
    ia = np.zeros(NF, np.float32); ia[X // 4] = 1.0
    ib = np.zeros(NF, np.float32); ib[Y // 4] = 1.0
    D = run(ia, ib, dh, dh, idesc(1, 1))
    print("A@%d B@%d -> nonzero at %s" % (X, Y, np.argwhere(D != 0)[:4].tolist()))
