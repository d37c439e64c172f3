"""Decode the record mudiff_debug_dump() returns after a conv_tc mbarrier wait timed out."""
NAMES = [(0, 'a_full', 16), (128, 'a_empty', 16), (256, 'b_full', 16), (384, 'b_empty', 16), (512, 'w_full', 1),
         (520, 'tfull', 2), (536, 'tempty', 2)]
ROLES = ['A-producer', 'B-producer', 'MMA0', 'MMA1', 'epi0', 'epi1', 'epi2', 'epi3']


def bar_name(off):
    for base, name, n in NAMES:
        if base <= off < base + 8 * n:
            return f"{name}[{(off - base) // 8}]"
    return {-2: 'bar.sync(bias)', -3: 'epilogue-body', -4: 'pre-tempty-arrive'}.get(off, f"?{off}")


def decode(d):
    d = [int(v) & 0xffffffff for v in d]
    out = []
    if not d[0]:
        return ['no timeout recorded']
    out.append(f"timeout: block {d[1]} warp {d[2]} lane {d[3]} waiting on {bar_name(d[4] - d[7])} parity {d[5]} (grid {d[6]})")
    blk = d[16:16 + 256]
    for base, name, n in NAMES:
        words = []
        for i in range(n):
            lo, hi = blk[(base + 8 * i) // 4], blk[(base + 8 * i) // 4 + 1]
            words.append(f"{hi:08x}:{lo:08x}")
        out.append(f"  {name:8s} " + ' '.join(words))
    out.append(f"  tmem_base {blk[552 // 4]:#x}")
    for w in range(8):
        r = blk[160 + 4 * w: 160 + 4 * w + 4]
        code = r[0] - (1 << 32) if r[0] >= (1 << 31) else r[0]
        tag = r[2] - (1 << 32) if r[2] >= (1 << 31) else r[2]
        state = {0: 'passed', 1: 'WAITING', 2: 'mark'}.get(r[3], str(r[3]))
        out.append(f"  warp {w} {ROLES[w]:10s}: {state:8s} {bar_name(code):18s} parity {r[1]} tag {tag}")
    return out


if __name__ == '__main__':
    import json, sys
    print('\n'.join(decode(json.load(open(sys.argv[1])))))
