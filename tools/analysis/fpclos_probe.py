"""Probe: restatement of FITPACK fpclos (s = 0, k = 3, closed curve) in plain float64 Python, compared bit for bit with
scipy.interpolate.splprep(per=1, s=0).  Exploratory."""
import math
import sys
import warnings

import numpy as np
from scipy.interpolate import splprep

sys.path.insert(0, "tests")
from helpers import golden  # noqa: E402


def fpbspl(t, k, x, l):          # l: 1-based interval index, t 0-based array
    h = [0.0] * (k + 2)
    hh = [0.0] * (k + 1)
    h[0] = 1.0
    for j in range(1, k + 1):
        for i in range(j):
            hh[i] = h[i]
        h[0] = 0.0
        for i in range(1, j + 1):
            li = l + i
            lj = li - j
            tli, tlj = t[li - 1], t[lj - 1]
            if tli == tlj:
                h[i] = 0.0
                continue
            f = hh[i - 1] / (tli - tlj)
            h[i - 1] = h[i - 1] + f * (tli - x)
            h[i] = f * (x - tlj)
    return h[:k + 1]


def fpgivs(piv, ww):
    store = abs(piv)
    if store >= ww:
        dd = store * math.sqrt(1.0 + (ww / piv) ** 2)
    else:
        dd = ww * math.sqrt(1.0 + (piv / ww) ** 2)
    return ww / dd, piv / dd, dd          # cos, sin, new ww


def fprota(c, s, a, b):
    return c * a - s * b, c * b + s * a   # new a, new b


def fpclos_s0_k3(u, x, y, variant=0):
    """u[0..M] (u[M] = 1), closed points x, y [M].  Returns c_x, c_y [M+3]."""
    k, k1 = 3, 4
    M = len(x)
    m = M + 1
    n = m + 2 * k
    t = np.empty(n)
    for i in range(2, m):               # t(j) = u(i), j = i + k, i = 2..m1
        t[i + k - 1] = u[i - 1]
    per = u[m - 1] - u[0]
    # boundary knots (fpclos: t(k1) = u(1), t(n-k) = u(m), then periodic extension)
    t[k1 - 1] = u[0]
    t[n - k - 1] = u[m - 1]
    for j in range(1, k + 1):
        i1 = n - k - 1 - j  # 0-based of t(n-k-j)
        t[k1 - 1 - j] = t[i1] - per
        t[n - k - 1 + j] = t[k1 - 1 + j] + per
    kk, kk1 = k - 1, k
    nk1 = n - k1
    n7 = nk1 - k
    n10 = n7 - kk
    a1 = np.zeros((n7 + 1, kk1 + 1))    # 1-based
    a2 = np.zeros((n7 + 1, kk + 1))
    z = np.zeros((2, n7 + 1))
    jper = 0
    l = k1
    for it in range(1, m):
        ui = u[it - 1]
        xi = [x[it - 1], y[it - 1]]
        while ui >= t[l]:               # t(l+1) 1-based = t[l]
            l += 1
        h = [0.0] + fpbspl(t, k, ui, l)  # 1-based h(1..k1)
        l5 = l - k1
        if l5 < n10:
            j = l5
            for i in range(1, kk1 + 1):
                j += 1
                piv = h[i]
                if piv == 0.0:
                    continue
                c, s, a1[j, 1] = fpgivs(piv, a1[j, 1])
                for d in range(2):
                    xi[d], z[d, j] = fprota(c, s, xi[d], z[d, j])
                if i == kk1:
                    break
                i2 = 1
                for i1 in range(i + 1, kk1 + 1):
                    i2 += 1
                    h[i1], a1[j, i2] = fprota(c, s, h[i1], a1[j, i2])
            continue
        if jper == 0:
            a2[:, :] = 0.0
            jk = n10 + 1
            for i in range(1, kk + 1):
                ik = jk
                for j in range(1, kk1 + 1):
                    if ik <= 0:
                        break
                    a2[ik, i] = a1[ik, j]
                    ik -= 1
                jk += 1
            jper = 1
        h1 = [0.0] * (kk1 + 2)
        h2 = [0.0] * (kk + 2)
        j = l5 - n10
        for i in range(1, kk1 + 1):
            j += 1
            l0 = j
            while True:
                l1 = l0 - kk
                if l1 <= 0:
                    h2[l0] = h2[l0] + h[i]
                    break
                if l1 <= n10:
                    h1[l1] = h[i]
                    break
                l0 = l1 - n10
        if n10 > 0:
            for j in range(1, n10 + 1):
                piv = h1[1]
                if piv == 0.0:
                    for i in range(1, kk + 1):
                        h1[i] = h1[i + 1]
                    h1[kk1] = 0.0
                    continue
                c, s, a1[j, 1] = fpgivs(piv, a1[j, 1])
                for d in range(2):
                    xi[d], z[d, j] = fprota(c, s, xi[d], z[d, j])
                for i in range(1, kk + 1):
                    h2[i], a2[j, i] = fprota(c, s, h2[i], a2[j, i])
                if j == n10:
                    break
                i2 = min(n10 - j, kk)
                i1 = 1
                for i in range(1, i2 + 1):
                    i1 = i + 1
                    h1[i1], a1[j, i1] = fprota(c, s, h1[i1], a1[j, i1])
                    h1[i] = h1[i1]
                h1[i1] = 0.0
        for j in range(1, kk + 1):
            ij = n10 + j
            if ij <= 0:
                continue
            piv = h2[j]
            if piv == 0.0:
                continue
            c, s, a2[ij, j] = fpgivs(piv, a2[ij, j])
            for d in range(2):
                xi[d], z[d, ij] = fprota(c, s, xi[d], z[d, ij])
            if j == kk:
                break
            for i in range(j + 1, kk + 1):
                h2[i], a2[ij, i] = fprota(c, s, h2[i], a2[ij, i])
    out = []
    for d in range(2):
        c = fpbacp(a1, a2, z[d], n7, kk, kk1)
        cc = np.empty(nk1)
        cc[:n7] = c[1:n7 + 1]
        for i in range(k):
            cc[n7 + i] = cc[i]
        out.append(cc)
    return t, out[0], out[1]


def fpbacp(a, b, z, n, k, k1):
    c = np.zeros(n + 2)
    n2 = n - k
    l = n
    for i in range(1, k + 1):
        store = z[l]
        j = k + 2 - i
        if i != 1:
            l0 = l
            for l1 in range(j, k + 1):
                l0 += 1
                store = store - c[l0] * b[l, l1]
        c[l] = store / b[l, j - 1]
        l -= 1
        if l == 0:
            return c
    for i in range(1, n2 + 1):
        store = z[i]
        l = n2
        for j in range(1, k + 1):
            l += 1
            store = store - c[l] * b[i, j]
        c[i] = store
    i = n2
    c[i] = c[i] / a[i, 1]
    if i == 1:
        return c
    for j in range(2, n2 + 1):
        i -= 1
        store = c[i]
        i1 = k
        if j <= k:
            i1 = j - 1
        l = i
        for l0 in range(1, i1 + 1):
            l += 1
            store = store - c[l] * a[i, l0 + 1]
        c[i] = store / a[i, 1]
    return c


if __name__ == "__main__":
    d = golden("cand_m579_n579")
    for b in range(3):
        p = d["points"][b]
        c = np.vstack([p, p[:1]])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            (t, (rx, ry), k), u = splprep([c[:, 0], c[:, 1]], s=0.0, k=3, per=1)
        tt, cx, cy = fpclos_s0_k3(u, p[:, 0], p[:, 1])
        print(b, "knots equal", np.array_equal(tt, t), "cx bit-equal", np.array_equal(cx, rx), "cy", np.array_equal(cy, ry),
              "max rel", np.max(np.abs(cx - rx)) / np.abs(rx).max(), "n differing", int((cx != rx).sum()), "of", len(rx))
