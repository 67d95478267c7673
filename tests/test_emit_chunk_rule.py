"""The chunk rule of the warp-per-read tuple emission (colord_b200/csrc/stage2_encode.cu: push_string, behind CLB_EMIT_WARP=1),
emulated lane by lane on the CPU and compared with the serial push of emit_read (encoder.cpp:1348-1443 run-length rules: a run of
>= 15 matches is one anchor tuple, a run of > 16 deletions one skip tuple, everything else one byte per symbol).  This checks the
RULE (run starts by ballot, first segment joins the pending run, complete runs sized and placed by a scan, last segment becomes
the pending run); the CUDA code that follows it has not run on a device yet."""
import random

T_INS, T_DEL, T_MATCH, T_SUB, T_ANCHOR, T_SKIP = 0, 1, 2, 3, 4, 5
CODE = {'A': 0, 'C': 1, 'G': 2, 'T': 3}

def sym_byte(s):
    if s == 'M': return T_MATCH << 4
    if s == 'D': return T_DEL << 4
    if s in 'XYZ': return (T_SUB << 4) + (ord(s) - ord('X'))
    return (T_INS << 4) + CODE[s]

def run_bytes(s, rep):          # TupleOut::run
    if s == 'M' and rep >= 15: return [(T_ANCHOR << 4) + (rep >> 24), (rep >> 16) & 255, (rep >> 8) & 255, rep & 255]
    if s == 'D' and rep > 16: return [(T_SKIP << 4) + (rep >> 24), (rep >> 16) & 255, (rep >> 8) & 255, rep & 255]
    return [sym_byte(s)] * rep

def serial(ps, pr, string):     # reference: push(c, 1) for every symbol; returns (bytes, ps, pr)
    out = []
    for c in string:
        if pr and ps == c: pr += 1
        else:
            if pr: out += run_bytes(ps, pr)
            ps, pr = c, 1
    return out, ps, pr

def warp(ps, pr, string):
    out = []
    k0 = 0
    while k0 < len(string):
        n = min(32, len(string) - k0)
        c = [string[k0 + t] if t < n else None for t in range(32)]
        # run starts: lane 0 compares with the pending run
        start = [False] * 32
        for t in range(n):
            prev = (ps if pr else None) if t == 0 else c[t - 1]
            start[t] = c[t] != prev
        mask = sum(1 << t for t in range(n) if start[t])
        valid = (1 << n) - 1
        if mask & 1:                       # the chunk opens a new run: the pending one is complete
            if pr: out += run_bytes(ps, pr)
            pr = 0
            lo = 0                         # first in-chunk segment starts at 0 and is an ordinary one
        else:
            # the first segment continues the pending run
            rest = mask & valid
            if rest == 0:                  # the whole chunk continues it
                pr += n; k0 += n; continue
            e0 = (rest & -rest).bit_length() - 1
            pr += e0
            out += run_bytes(ps, pr); pr = 0
            lo = e0
        # segments from lo on live inside the chunk; the last one becomes the pending run
        m2 = mask & valid & ~((1 << lo) - 1)          # bit lo is set
        last_b = m2.bit_length() - 1
        # per lane: segment start b, end e, complete?
        contrib = [0] * 32; info = [None] * 32
        for t in range(lo, n):
            below = m2 & ((2 << t) - 1)
            b = below.bit_length() - 1
            above = m2 & ~((2 << t) - 1)
            e = ((above & -above).bit_length() - 1) if above else n
            complete = b != last_b
            L = e - b
            if complete:
                four = (c[b] == 'M' and L >= 15) or (c[b] == 'D' and L > 16)
                contrib[t] = (4 if t == b else 0) if four else 1
                info[t] = (b, L, four)
        # exclusive scan of contrib
        off = [0] * 32; acc = 0
        for t in range(32): off[t] = acc; acc += contrib[t]
        buf = [None] * acc
        for t in range(lo, n):
            if info[t] is None: continue
            b, L, four = info[t]
            if four:
                if t == b:
                    buf[off[t]:off[t] + 4] = run_bytes(c[b], L)
            else:
                buf[off[t]] = sym_byte(c[t])
        assert all(x is not None for x in buf)
        out += buf
        ps, pr = c[last_b], n - last_b
        k0 += n
    return out, ps, pr

def test_warp_chunk_rule_equals_serial_push():
  random.seed(1)
  for trial in range(4000):
      n = random.choice([0, 1, 2, 5, 31, 32, 33, 64, 70, 100, random.randint(0, 200)])
      alpha = random.choice(["MD", "MDACGTXYZ", "M", "MMMMMMMD", "DDDDDDDM", "ACGT", "MDX"])
      string = [random.choice(alpha) for _ in range(n)]
      ps = random.choice("MDACGTXYZ"); pr = random.choice([0, 0, 1, 3, 14, 15, 16, 17, 500])
      a = serial(ps, pr, string); b = warp(ps, pr, string)
      # the pending symbol is irrelevant while pr == 0
      if a[0] != b[0] or a[2] != b[2] or (a[2] and a[1] != b[1]):
          raise AssertionError(("mismatch", trial, ps, pr, "".join(string), a, b))
