"""Radix-2 NTT over F_r with the reference's orderings (oracle; test infrastructure only).

`taceo_ark_algebra::fft::Domain::{ifft_in_to_out, fft_out_to_in}` (call sites
co-groth16/src/groth16/reduction.rs:141-175): "in" = natural order, "out" = bit-reversed order;
the inverse includes the 1/n scaling.  Plonk uses natural-in/natural-out `domain.fft/ifft`
with a hand-set group_gen (co-plonk/src/types.rs:76-100).
"""
from .fields import inv


def bit_reverse_perm(v):
    """fft::bit_reverse (reduction.rs:58,328): out[rev(i)] = in[i]."""
    n = len(v)
    lg = n.bit_length() - 1
    assert 1 << lg == n
    out = list(v)
    for i in range(n):
        j = int(format(i, "0%db" % lg)[::-1], 2) if lg else 0
        if i < j:
            out[i], out[j] = out[j], out[i]
    return out


def _dif(v, w, r):
    """Gentleman-Sande: natural in -> bit-reversed out; computes sum_j v_j w^{ij} at position rev(i)."""
    n = len(v)
    a = list(v)
    m = n // 2
    wm = w
    while m >= 1:
        for k in range(0, n, 2 * m):
            t = 1
            for j in range(m):
                x, y = a[k + j], a[k + j + m]
                a[k + j] = (x + y) % r
                a[k + j + m] = (x - y) * t % r
                t = t * wm % r
        wm = wm * wm % r
        m //= 2
    return a


def _dit(v, w, r):
    """Cooley-Tukey: bit-reversed in -> natural out."""
    n = len(v)
    a = list(v)
    lg = n.bit_length() - 1
    ws = [w]
    for _ in range(lg - 1):
        ws.append(ws[-1] * ws[-1] % r)
    m = 1
    s = lg - 1
    while m < n:
        wm = ws[s] if lg else 1
        for k in range(0, n, 2 * m):
            t = 1
            for j in range(m):
                x, y = a[k + j], a[k + j + m] * t % r
                a[k + j] = (x + y) % r
                a[k + j + m] = (x - y) % r
                t = t * wm % r
        m *= 2
        s -= 1
    return a


def fft(v, gen, r):
    """natural in, natural out: out[i] = sum_j v[j] gen^{ij}."""
    return bit_reverse_perm(_dif(v, gen, r))


def ifft(v, gen, r):
    n = len(v)
    ni = inv(n % r, r)
    return [x * ni % r for x in fft(v, inv(gen, r), r)]


def ifft_in_to_out(v, gen, r):
    """natural in -> bit-reversed out, scaled by 1/n."""
    n = len(v)
    ni = inv(n % r, r)
    return [x * ni % r for x in _dif(v, inv(gen, r), r)]


def fft_out_to_in(v, gen, r):
    """bit-reversed in -> natural out."""
    return _dit(v, gen, r)
