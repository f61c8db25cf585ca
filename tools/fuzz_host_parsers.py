#!/usr/bin/env python
"""Mutation fuzz of the host-side file parsers (csrc/bq2_cache.c: MD2 / MD3 file images, MD2 / MD3 cache files) and
the neighbour builders (csrc/bq2_host.c) under AddressSanitizer + UBSan.  The parsers take untrusted bytes; every
input sits in an exact-size heap block so any over-read is seen.  No GPU involved.

    cd berserkerquake2_b200/csrc
    gcc -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -fPIC -shared -ffp-contract=off \
        -I../../include -I. bq2_cache.c bq2_host.c bq2_synth.c -lm -o /tmp/libhost_asan.so
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \
        python tools/fuzz_host_parsers.py <seed> <iterations> [/tmp/libhost_asan.so]

Round 1: seeds 1-3 x 8,000 iterations clean (about a third of the mutated images are accepted and converted) after the
parsers were changed to read multi-byte fields through memcpy and md3_mesh_at to refuse a negative ofs_meshes."""
import ctypes as C, numpy as np, struct, sys
lib = C.CDLL(sys.argv[3] if len(sys.argv) > 3 else "/tmp/libhost_asan.so")
vp = C.c_void_p
def P(a): return a.ctypes.data_as(vp)
lib.bq2_synth_md2_bytes.restype = C.c_size_t
lib.bq2_synth_md2_bytes.argtypes = [C.c_int32]*3
lib.bq2_synth_md2.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_uint32, vp]
lib.bq2_host_md2_from_file.argtypes = [vp, C.c_size_t, C.c_float, C.c_int, vp, C.c_size_t, vp]
lib.bq2_md3_file_info.argtypes = [vp, C.c_size_t, vp, vp, vp, vp]
lib.bq2_host_md3_mesh_from_file.argtypes = [vp, C.c_size_t, C.c_int32, C.c_float, C.c_int, vp, vp, vp, vp]
lib.bq2_md2_cache_read.argtypes = [vp, C.c_size_t, C.c_int, C.c_float, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
lib.bq2_md2_cache_bytes.restype = C.c_size_t
lib.bq2_md2_cache_bytes.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_int32]
lib.bq2_md2_cache_write.argtypes = [vp, C.c_size_t, C.c_int, C.c_float, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]
lib.bq2_md3_cache_read.argtypes = [vp, C.c_size_t, C.c_uint32, vp, C.c_int32, C.c_int32]
lib.bq2_host_md2_neighbours.argtypes = [vp, C.c_int32, vp]
lib.bq2_host_md3_neighbours.argtypes = [vp, C.c_int32, vp]
rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000

S, R, F = 6, 5, 3
nb = lib.bq2_synth_md2_bytes(S, R, F)
md2 = np.zeros(nb, np.uint8); assert lib.bq2_synth_md2(S, R, F, 7, P(md2)) == 0

def build_md3(F=2, V=5, T=4, M=2):
    frames = b"".join(struct.pack("<10f16s", -1, -2, -3, 1, 2, 3, 1.0, 2.0, 3.0, 9.0, b"test") for _ in range(F))
    body = b""
    for m in range(M):
        skins = struct.pack("<64si", b"models/x/skin.tga", 0)
        tris = rng.randint(0, V, (T, 3)).astype("<u4").tobytes()
        tcs = rng.rand(V, 2).astype("<f4").tobytes()
        verts = rng.randint(-3000, 3000, (F, V, 4)).astype("<i2")
        ofs_tris = 108 + len(skins); ofs_tcs = ofs_tris + len(tris); ofs_verts = ofs_tcs + len(tcs)
        size = ofs_verts + verts.nbytes
        body += struct.pack("<4s64s11i", b"IDP3", b"mesh", 0, F, 1, V, T, ofs_tris, 108, ofs_tcs, ofs_verts, size, 0)[:108]
        body += skins + tris + tcs + verts.tobytes()
    hdr = struct.pack("<ii64s9i", 0x33504449, 15, b"test.md3", 0, F, 0, M, 0, 108, 108 + len(frames), 108 + len(frames), 108 + len(frames) + len(body))
    return np.frombuffer(hdr + frames + body, np.uint8).copy()
md3 = build_md3()

def mutate(a, header_ints):
    b = a.copy()
    k = rng.randint(1, 4)
    for _ in range(k):
        mode = rng.randint(4)
        if mode == 0:                       # a header int gets an interesting value
            i = header_ints[rng.randint(len(header_ints))]
            v = int(rng.choice([0, -1, 1, 2, 0x7fffffff, -0x80000000, 0x40000000, len(a), len(a) - 1, len(a) + 1, rng.randint(-70000, 70000)]))
            if i + 4 <= len(b): b[i:i + 4] = np.frombuffer(struct.pack("<i", v), np.uint8)
        elif mode == 1:                     # random byte flips anywhere
            for _ in range(rng.randint(1, 8)):
                if len(b): b[rng.randint(len(b))] = rng.randint(256)
        elif mode == 2:                     # truncate
            b = b[:rng.randint(0, len(b) + 1)].copy()
        else:                               # small delta on a header int
            i = header_ints[rng.randint(len(header_ints))]
            v = struct.unpack("<i", b[i:i + 4].tobytes())[0] + int(rng.randint(-5, 6)) if i + 4 <= len(b) else 0
            if i + 4 <= len(b): b[i:i + 4] = np.frombuffer(struct.pack("<i", max(-2**31, min(2**31 - 1, v))), np.uint8)
    return b

md2_ints = list(range(0, 68, 4))
md3_hdr_ints = [4] + list(range(72, 108, 4))
first_mesh = 108 + 2 * 56
md3_ints = md3_hdr_ints + list(range(first_mesh + 68, first_mesh + 108, 4))
ok2 = ok3 = 0
for it in range(N):
    f = mutate(md2, md2_ints)
    exact = np.empty(max(len(f), 1), np.uint8); exact[:len(f)] = f            # exact-size heap block: ASan sees any over-read
    if len(f) >= 68:
        h = f[:68].view("<i4")
        num_st = int(h[7])
        st = np.empty(max(2 * max(num_st, 0), 2) if 0 <= num_st <= len(f) else 2, np.float32)
    else:
        st = np.empty(2, np.float32)
    out = np.empty(max(len(f), 1), np.uint8)
    rc = lib.bq2_host_md2_from_file(P(exact), len(f), 1.0, int(rng.randint(2)), P(out), len(f), P(st))
    ok2 += rc == 0
    if rc == 0:                                                               # what a caller does next: neighbours from the image
        h = out[:68].view("<i4"); T = int(h[8]); o = int(h[13])
        if o < 68 or o + 12 * T > len(out):                                     # (tris overlapped the header: a caller re-validates, as bq2_upload_md2 does)
            continue
        tris = out[o:o + 12 * T].copy().view(np.int16)
        nbt = np.empty((T, 3), np.int32)
        lib.bq2_host_md2_neighbours(P(tris), T, P(nbt))
    g = mutate(md3, md3_ints)
    exact3 = np.empty(max(len(g), 1), np.uint8); exact3[:len(g)] = g
    nF, nM = C.c_int32(), C.c_int32(); nv = np.zeros(32, np.int32); nt = np.zeros(32, np.int32)
    rc = lib.bq2_md3_file_info(P(exact3), len(g), C.byref(nF), C.byref(nM), P(nv), P(nt))
    if rc == 0:
        ok3 += 1
        for m in range(nM.value):
            V, T, Fm = int(nv[m]), int(nt[m]), nF.value
            vx = np.empty(16 * Fm * V, np.uint8); ix = np.empty(3 * T, np.uint16); stm = np.empty(2 * V, np.float32); ft = np.empty(3 * Fm, np.float32)
            r2 = lib.bq2_host_md3_mesh_from_file(P(exact3), len(g), m, 1.0, int(rng.randint(2)), P(ft), P(vx), P(ix), P(stm))
            if r2 == 0:
                nbt = np.empty((T, 3), np.int32); lib.bq2_host_md3_neighbours(P(ix), T, P(nbt))
    # cache files: random bytes of random length against fixed dimensions
    T, rows, Fc = 16, 10, 3
    for smooth in (0, 1):
        need = lib.bq2_md2_cache_bytes(smooth, T, rows, Fc)
        cb = rng.randint(0, 256, rng.choice([need, need - 1, need + 1, rng.randint(0, need + 8)])).astype(np.uint8)
        cb = np.ascontiguousarray(cb) if len(cb) else np.empty(1, np.uint8)
        why = C.c_int32()
        lib.bq2_md2_cache_read(P(cb), len(cb) if len(cb) > 1 or need <= 1 else len(cb), smooth, 0.7, T, rows, Fc, P(np.empty((T, 3), np.int32)), P(np.empty(rows * Fc, np.uint8)),
                               P(np.empty(rows * Fc, np.uint8)), P(np.empty(rows * Fc, np.uint8)), C.byref(why))
    c3 = rng.randint(0, 256, rng.randint(0, 4 + 3 * 2 * 5 + 4)).astype(np.uint8)
    c3 = c3 if len(c3) else np.empty(1, np.uint8)
    lib.bq2_md3_cache_read(P(c3), len(c3), 1234, P(np.empty(16 * 2 * 5, np.uint8)), 2, 5)
print("iterations", N, "accepted md2", ok2, "accepted md3", ok3)
