"""Test helper: executes a ``front.FrontSpec`` in numpy exactly the way ``csrc/front_tc.cuh`` walks it.

Decodes the weight chunks the kernel streams (head | FP16 hi image | FP16 lo image, canonical core-matrix order),
applies the per-level scale, the children's bias and saturation in the consumer, and the kernel's term order.  Agreement
with ``plan_interp`` on ops 0..2 proves tables, chunk layout and folding on the CPU; the arithmetic itself (FP16 pieces on
tcgen05) is checked on the GPU."""
import numpy as np

from pyfaceanalysis_b200 import front as fr


def _decode(chunk, N):
    head = np.frombuffer(chunk[:fr.HEAD_BYTES], dtype=np.float32).astype(np.float64)
    img = np.frombuffer(chunk[fr.HEAD_BYTES:], dtype=np.float16).astype(np.float64)
    half = fr.CHUNK_TERMS * N
    def canon(a):     # [k/8][n/8][n%8][k%8] -> (32, N)
        return a.reshape(fr.CHUNK_TERMS // 8, N // 8, 8, 8).transpose(0, 3, 1, 2).reshape(fr.CHUNK_TERMS, N)
    return head, canon(img[:half]) + canon(img[half:2 * half])


def run_front(f, x):
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    img = x.reshape(n, f.img_h, f.img_w)
    out = np.zeros((n, f.out_dim))
    cb = [f.chunk_bytes(k) for k in range(3)]
    for k in range(f.n_sub):
        X, Y = f.pair_xy[k // 2]
        sub = f.wimg[k * f.sub_bytes:(k + 1) * f.sub_bytes]
        pos = 0

        def take(level):
            nonlocal pos
            c = sub[pos:pos + cb[level]]
            pos += cb[level]
            return _decode(c, f.nn[level])
        acc0 = []
        for i in range(4):
            dy, dx = int(f.l0_off[k, i]) & 0xff, int(f.l0_off[k, i]) >> 8
            px = img[:, Y + dy:Y + dy + 4, X + dx:X + dx + 4].reshape(n, 16)
            head, W = take(0)
            A = np.concatenate([px, np.abs(px - head[:16]) ** f.pexp[0]], axis=1)
            acc0.append(A @ W)

        def join(level, NP, kids, s_child, clip_child):
            parts = [take(level) for _ in range(f.nch[level])]
            head = parts[0][0]
            W = np.concatenate([p[1] for p in parts], axis=0)
            A = np.zeros((n, 4 * NP))
            for c in range(2):          # one warp per child: its [identity | power] sequence fills its half of every chunk
                y = np.clip(kids[c][:, :NP] * s_child + head[c * NP:(c + 1) * NP], *clip_child)
                seq = np.concatenate([y, np.abs(y - head[2 * NP + c * NP:2 * NP + (c + 1) * NP]) ** f.pexp[level]], axis=1)
                for s_ in range(2 * NP):
                    A[:, fr.join_column(NP, c, s_)] = seq[:, s_]
            return A @ W, head
        acc1 = [join(1, f.np1, acc0[2 * h:2 * h + 2], f.scale[0], f.clip[0])[0] for h in range(2)]
        acc2, head2 = join(2, f.np2, acc1, f.scale[1], f.clip[1])
        y = np.clip(acc2[:, :32] * f.scale[2] + head2[4 * f.np2:4 * f.np2 + acc2[:, :32].shape[1]], *f.clip[2])
        out[:, f.out_col[k]:f.out_col[k] + f.nv[2]] = y[:, :f.nv[2]]
        assert pos == f.sub_bytes
    return out
