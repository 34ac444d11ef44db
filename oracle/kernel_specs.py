"""TEST INFRASTRUCTURE — executable specification of every C-ABI kernel.

One torch (CPU, fp32 or fp64) function per entry point of
``include/chgnet_b200.h``, same argument order and caller-allocated outputs.
The reverse kernels are written as EXPLICIT formulas (no autograd) — they are
the maths the CUDA kernels transcribe; ``tests/test_engine_spec.py`` checks the
whole chain against ``oracle/chgnet_oracle.py`` (autograd, fp64), and the
``-m gpu`` tests check each CUDA kernel against the function of the same name
here.  Never imported by the product path.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor


def _sig(x):
    return torch.sigmoid(x)


def _silu(x):
    return x * _sig(x)


def _dsilu(x):
    s = _sig(x)
    return s * (1 + x * (1 - s))


def _ln_fwd(p, g, b, eps=1e-5):
    mu = p.mean(dim=1, keepdim=True)
    var = ((p - mu) ** 2).mean(dim=1, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + eps)
    xhat = (p - mu) * rstd
    return xhat * g + b, xhat, rstd


def _ln_bwd(gy, xhat, rstd, g):
    gx = gy * g
    return rstd * (gx - gx.mean(dim=1, keepdim=True) - xhat * (gx * xhat).mean(dim=1, keepdim=True))


def _gate_fwd(p, ln):
    """p [M,128] -> out [M,64] and the pieces the reverse needs."""
    pc, pg = p[:, :64], p[:, 64:]
    if ln is not None:
        y1, xh1, r1 = _ln_fwd(pc, ln[0], ln[1])
        y2, xh2, r2 = _ln_fwd(pg, ln[2], ln[3])
    else:
        y1, y2, xh1, xh2, r1, r2 = pc, pg, None, None, None, None
    core, gate = _silu(y1), _sig(y2)
    return core * gate, (y1, y2, core, gate, xh1, xh2, r1, r2)


def _gate_bwd(g_out, saved, ln, g_ln=None):
    y1, y2, core, gate, xh1, xh2, r1, r2 = saved
    gy1 = g_out * gate * _dsilu(y1)
    gy2 = g_out * core * gate * (1 - gate)
    if ln is not None and g_ln is not None:  # training: d/d(gamma1, beta1, gamma2, beta2), accumulated
        g_ln += torch.stack([(gy1 * xh1).sum(0), gy1.sum(0), (gy2 * xh2).sum(0), gy2.sum(0)]).reshape(-1).to(g_ln.dtype)
    if ln is not None:
        gy1 = _ln_bwd(gy1, xh1, r1, ln[0])
        gy2 = _ln_bwd(gy2, xh2, r2, ln[2])
    return torch.cat([gy1, gy2], dim=1)


def _rbf(d, freq, rc, p):
    """basis [M,R] and d(basis)/dd [M,R]  (SURVEY.md appendix B)."""
    d = d[:, None]
    x = d / rc
    nrm = math.sqrt(2.0 / rc)
    s, c = torch.sin(freq * x), torch.cos(freq * x)
    if p != 0:
        a, b, cc = -(p + 1) * (p + 2) / 2, p * (p + 2), -p * (p + 1) / 2
        env = 1 + a * x**p + b * x ** (p + 1) + cc * x ** (p + 2)
        denv = (a * p * x ** (p - 1) + b * (p + 1) * x**p + cc * (p + 2) * x ** (p + 1)) / rc
        inside = x < 1
        env = torch.where(inside, env, torch.zeros_like(env))
        denv = torch.where(inside, denv, torch.zeros_like(denv))
    else:
        env, denv = torch.ones_like(x), torch.zeros_like(x)
    raw = nrm * s / d
    draw = nrm * ((freq / rc) * c / d - s / d**2)
    return raw * env, draw * env + raw * denv


def _rbf_dfreq(d, freq, rc, p):
    """d(basis_k)/d(freq_k) [M,R]: sqrt(2/rc) cos(w d/rc)/rc * env."""
    d = d[:, None]
    x = d / rc
    nrm = math.sqrt(2.0 / rc)
    if p != 0:
        a, b, cc = -(p + 1) * (p + 2) / 2, p * (p + 2), -p * (p + 1) / 2
        env = 1 + a * x**p + b * x ** (p + 1) + cc * x ** (p + 2)
        env = torch.where(x < 1, env, torch.zeros_like(env))
    else:
        env = torch.ones_like(x)
    return nrm * torch.cos(freq * x) / rc * env



# ---- second-order pieces (force / stress losses: d/dtheta of T = <dE/dr, rdot>) ------------------
def _d2silu(x):
    s = _sig(x)
    return s * (1 - s) * (2 + x * (1 - 2 * s))


def _dsig(x):
    s = _sig(x)
    return s * (1 - s)


def _d2sig(x):
    s = _sig(x)
    return s * (1 - s) * (1 - 2 * s)


def _ln_tan(pd, xhat, rstd):
    """tangent of xhat = (p - mean) rstd along pd"""
    return rstd * (pd - pd.mean(dim=1, keepdim=True) - xhat * (xhat * pd).mean(dim=1, keepdim=True))


def _gate_tan(p, pd, ln):
    """o = silu(LN1 p_core) sigmoid(LN2 p_gate) and its tangent along pd; cache for _gate_bwd2."""
    pc, pg, pdc, pdg = p[:, :64], p[:, 64:], pd[:, :64], pd[:, 64:]
    if ln is not None:
        y1, xh1, r1 = _ln_fwd(pc, ln[0], ln[1])
        y2, xh2, r2 = _ln_fwd(pg, ln[2], ln[3])
        xd1, xd2 = _ln_tan(pdc, xh1, r1), _ln_tan(pdg, xh2, r2)
        yd1, yd2 = ln[0] * xd1, ln[2] * xd2
    else:
        y1, y2, yd1, yd2 = pc, pg, pdc, pdg
        xh1 = xh2 = r1 = r2 = xd1 = xd2 = None
    o = _silu(y1) * _sig(y2)
    od = _dsilu(y1) * _sig(y2) * yd1 + _silu(y1) * _dsig(y2) * yd2
    return o, od, (y1, y2, yd1, yd2, xh1, xh2, r1, r2, xd1, xd2, pdc, pdg)


def _gate_bwd2(go, a, cache, ln, g_ln=None):
    """u = d/dp [ <go, o(p)> + <a, Do(p)[pd]> ] with go, a, pd held fixed; the LayerNorm affine
    gradients of the same scalar are accumulated into g_ln ([4][64] flat: g1, b1, g2, b2)."""
    y1, y2, yd1, yd2, xh1, xh2, r1, r2, xd1, xd2, pdc, pdg = cache
    s, ds, d2s = _silu(y1), _dsilu(y1), _d2silu(y1)
    t, dt, d2t = _sig(y2), _dsig(y2), _d2sig(y2)
    gy1 = go * t * ds + a * (d2s * t * yd1 + ds * dt * yd2)
    gy2 = go * s * dt + a * (ds * dt * yd1 + s * d2t * yd2)
    k1, k2 = a * ds * t, a * s * dt  # d/d(ydot)
    if ln is None:
        return torch.cat([gy1, gy2], dim=1)

    def branch(gy, kap, gamma, xh, r, xd, pd):
        kk = kap * gamma
        v = gy * gamma + r * (-kk * (xh * pd).mean(dim=1, keepdim=True) - (kk * xh).sum(dim=1, keepdim=True) * pd / 64)
        q = r * (v - v.mean(dim=1, keepdim=True) - xh * (v * xh).mean(dim=1, keepdim=True))
        return q - r * xh * (kk * xd).sum(dim=1, keepdim=True) / 64

    if g_ln is not None:
        g_ln += torch.stack([(gy1 * xh1 + k1 * xd1).sum(0), gy1.sum(0), (gy2 * xh2 + k2 * xd2).sum(0), gy2.sum(0)]).reshape(-1).to(g_ln.dtype)
    return torch.cat([branch(gy1, k1, ln[0], xh1, r1, xd1, pdc), branch(gy2, k2, ln[2], xh2, r2, xd2, pdg)], dim=1)


def _rbf_d_dfreq(d, freq, rc, p):
    """d/dfreq_k of (d basis_k / dd)  [M,R]  (mixed second derivative)."""
    d = d[:, None]
    x = d / rc
    nrm = math.sqrt(2.0 / rc)
    sn, cs = torch.sin(freq * x), torch.cos(freq * x)
    if p != 0:
        a, b, cc = -(p + 1) * (p + 2) / 2, p * (p + 2), -p * (p + 1) / 2
        env = 1 + a * x**p + b * x ** (p + 1) + cc * x ** (p + 2)
        denv = (a * p * x ** (p - 1) + b * (p + 1) * x**p + cc * (p + 2) * x ** (p + 1)) / rc
        inside = x < 1
        env = torch.where(inside, env, torch.zeros_like(env))
        denv = torch.where(inside, denv, torch.zeros_like(denv))
    else:
        env, denv = torch.ones_like(x), torch.zeros_like(x)
    # raw = nrm sin(w x)/d ; draw = nrm [ (w/rc) cos(w x)/d - sin(w x)/d^2 ]
    draw_dw = nrm * ((1 / rc) * cs / d - (freq / rc) * x * sn / d - x * cs / d**2)
    raw_dw = nrm * x * cs / d
    return draw_dw * env + raw_dw * denv


class SpecKernels:
    """Drop-in for ``chgnet_b200._lib.CudaKernels`` in CPU tests."""

    name = "spec"
    launches = 0

    # ---- K0
    def embed_atoms(self, z, emb, x):
        x.copy_(emb[z.long() - 1])

    # ---- K1a
    def edge_geometry(self, frac, lattice, owner, center, nbr, image, rvec, dist, rhat):
        L = lattice.view(-1, 3, 3)
        cart = torch.einsum("ni,nij->nj", frac, L[owner.long()])
        c, n = center.long(), nbr.long()
        Le = L[owner.long()[c]]
        r = cart[c] - (cart[n] + torch.einsum("ei,eij->ej", image, Le))
        d = torch.linalg.norm(r, dim=1)
        rvec.copy_(r), dist.copy_(d), rhat.copy_(r / d[:, None])

    # ---- K1b
    def bond_basis_embed(self, dist, u2d, freq_ag, freq_bg, rc_ag, rc_bg, p, w3t, e0, wag, wbg, basis_out=None):
        du = dist[u2d.long()]
        bag, _ = _rbf(du, freq_ag, rc_ag, p)
        bbg, _ = _rbf(du, freq_bg, rc_bg, p)
        e0.copy_(bag @ w3t[0]), wag.copy_(bag @ w3t[1]), wbg.copy_(bbg @ w3t[2])
        if basis_out is not None:  # training: [Eu][64] = ag basis in columns 0.., bg basis in columns 32..
            R = freq_ag.shape[0]
            basis_out.zero_()
            basis_out[:, :R] = bag
            basis_out[:, 32 : 32 + R] = bbg

    def bond_basis_bwd(self, dist, u2d, freq_ag, freq_bg, rc_ag, rc_bg, p, w3, g_e0, g_wag, g_wbg, g_dist, g_freq=None):
        du = dist[u2d.long()]
        _, dag = _rbf(du, freq_ag, rc_ag, p)
        _, dbg = _rbf(du, freq_bg, rc_bg, p)
        gb_ag = g_e0 @ w3[0] + g_wag @ w3[1]
        gb_bg = g_wbg @ w3[2]
        g_dist.copy_((gb_ag * dag).sum(dim=1) + (gb_bg * dbg).sum(dim=1))
        if g_freq is not None:  # training: d/d(freq_ag), d/d(freq_bg) as fp64 [2][R], accumulated
            g_freq[0] += (gb_ag * _rbf_dfreq(du, freq_ag, rc_ag, p)).sum(dim=0).to(g_freq.dtype)
            g_freq[1] += (gb_bg * _rbf_dfreq(du, freq_bg, rc_bg, p)).sum(dim=0).to(g_freq.dtype)

    # ---- K2
    def angle_basis_embed(self, rhat, ang_di, ang_dj, freq, wt, a0, basis_out=None):
        u = (rhat[ang_di.long()] * rhat[ang_dj.long()]).sum(dim=1) * (1 - 1e-6)
        th = torch.acos(u)
        arg = th[:, None] * freq[None, :]
        f = torch.cat([torch.full_like(th[:, None], 1 / math.sqrt(2.0)), torch.sin(arg), torch.cos(arg)], dim=1)
        a0.copy_((f / math.sqrt(math.pi)) @ wt)
        if basis_out is not None:  # training: [A][64], Fourier basis in columns 0..2F, zero padding after
            basis_out.zero_()
            basis_out[:, : f.shape[1]] = f / math.sqrt(math.pi)

    def angle_basis_bwd(self, rhat, ang_di, ang_dj, freq, w, g_a0, g_rhat, g_freq=None):
        ri, rj = rhat[ang_di.long()], rhat[ang_dj.long()]
        u = (ri * rj).sum(dim=1) * (1 - 1e-6)
        th = torch.acos(u)
        arg = th[:, None] * freq[None, :]
        nf = freq.shape[0]
        gf = (g_a0 @ w) / math.sqrt(math.pi)  # [A, 2F+1]
        g_th = (gf[:, 1 : 1 + nf] * torch.cos(arg) * freq).sum(dim=1) - (gf[:, 1 + nf :] * torch.sin(arg) * freq).sum(dim=1)
        g_u = -g_th / torch.sqrt(1 - u * u) * (1 - 1e-6)
        if g_freq is not None:  # training: d/d(freq) fp64 [F], accumulated
            g_freq += ((gf[:, 1 : 1 + nf] * torch.cos(arg) - gf[:, 1 + nf :] * torch.sin(arg)) * th[:, None]).sum(dim=0).to(g_freq.dtype)
        if g_rhat is None:
            return
        g_rhat.index_add_(0, ang_di.long(), (g_u[:, None] * rj).to(g_rhat.dtype))
        g_rhat.index_add_(0, ang_dj.long(), (g_u[:, None] * ri).to(g_rhat.dtype))

    # ---- dense
    def linear(self, x, wt, bias, residual, y, x_rows=None, y_rows=None):
        out = (x if x_rows is None else x[x_rows.long()]) @ wt
        if bias is not None:
            out = out + bias
        if y_rows is None:
            if residual is not None:
                out = out + residual
            y.copy_(out)
        else:
            if residual is not None:
                out = out + residual[y_rows.long()]
            y[y_rows.long()] = out

    def gather_rows(self, src, idx, dst):
        dst.copy_(src[idx.long()])

    def scatter_rows(self, src, idx, dst):
        dst[idx.long()] = src

    # ---- K4
    def _atom_pre(self, pcn, pe, center, nbr, d2u):
        return pcn[center.long(), :128] + pe[d2u.long()] + pcn[nbr.long(), 128:]

    def atom_conv_fwd(self, pcn, pe, wag, center, nbr, d2u, w2t, b2, ln, msg, save_p, save_pre=None):
        if save_pre is not None:
            save_pre.copy_(self._atom_pre(pcn, pe, center, nbr, d2u))
        h = _silu(self._atom_pre(pcn, pe, center, nbr, d2u))
        p = torch.cat([h[:, :64] @ w2t[:, :64], h[:, 64:] @ w2t[:, 64:]], dim=1) + b2
        out, _ = _gate_fwd(p, ln)
        msg.copy_(out * wag[d2u.long()])
        if save_p is not None:
            save_p.copy_(p)

    def atom_conv_bwd(self, pcn, pe, wag, center, nbr, d2u, save_p, g_agg, w2, ln, g_pre, g_w, g_p_out=None, g_ln=None):
        pre = self._atom_pre(pcn, pe, center, nbr, d2u)
        out, saved = _gate_fwd(save_p, ln)
        g_msg = g_agg[center.long()]
        g_w.copy_(g_msg * out)
        g_p = _gate_bwd(g_msg * wag[d2u.long()], saved, ln, g_ln)
        if g_p_out is not None:
            g_p_out.copy_(g_p)
        g_h = torch.cat([g_p[:, :64] @ w2[:64], g_p[:, 64:] @ w2[64:]], dim=1)
        g_pre.copy_(g_h * _dsilu(pre))

    # ---- K4f / K5f: message + aggregation fused (chg_atom_conv_fused / chg_bond_conv_fused)
    def atom_conv_fused(self, pcn, pe, wag, center, nbr, d2u, ptr_c, w2t, b2, ln, agg, save_p):
        msg = torch.empty(center.shape[0], 64, dtype=agg.dtype)
        SpecKernels.atom_conv_fwd(self, pcn, pe, wag, center, nbr, d2u, w2t, b2, ln, msg, save_p)  # unbound: not re-recorded
        SpecKernels.segment_sum(self, msg, None, ptr_c, 0, agg)

    def bond_conv_fused(self, pij, px, pa, wbg, ang_atom, ang_i, ang_j, ptr_i, w2t, b2, ln, agg, save_pre, save_p):
        upd = torch.empty(ang_i.shape[0], 64, dtype=agg.dtype)
        SpecKernels.bond_conv_fwd(self, pij, px, pa, wbg, ang_atom, ang_i, ang_j, w2t, b2, ln, upd, save_pre, save_p)
        SpecKernels.segment_sum(self, upd, None, ptr_i, 0, agg)

    # ---- K4s
    def segment_sum(self, data, perm, ptr, accumulate, out):
        n_rows = ptr.shape[0] - 1
        counts = (ptr[1:] - ptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(n_rows), counts)
        total = int(ptr[-1])
        src = (data if perm is None else data[perm.long()])[:total]
        acc = torch.zeros_like(out).index_add_(0, rows, src)
        if accumulate:
            out.add_(acc)
        else:
            out.copy_(acc)

    # ---- K5
    def _bond_pre(self, pij, px, pa, ang_atom, ang_i, ang_j):
        return pij[ang_i.long(), :128] + pij[ang_j.long(), 128:] + px[ang_atom.long()] + pa

    def bond_conv_fwd(self, pij, px, pa, wbg, ang_atom, ang_i, ang_j, w2t, b2, ln, upd, save_pre, save_p):
        pre = self._bond_pre(pij, px, pa, ang_atom, ang_i, ang_j)
        h = _silu(pre)
        p = torch.cat([h[:, :64] @ w2t[:, :64], h[:, 64:] @ w2t[:, 64:]], dim=1) + b2
        out, _ = _gate_fwd(p, ln)
        upd.copy_(out * wbg[ang_i.long()] * wbg[ang_j.long()])
        if save_pre is not None:
            save_pre.copy_(pre)
        if save_p is not None:
            save_p.copy_(p)

    def bond_conv_bwd(self, save_pre, save_p, wbg, ang_i, ang_j, g_agg, w2, ln, g_pre, gw_i, gw_j, g_p_out=None, g_ln=None):
        out, saved = _gate_fwd(save_p, ln)
        wi, wj = wbg[ang_i.long()], wbg[ang_j.long()]
        g_upd = g_agg[ang_i.long()]
        gw_i.copy_(g_upd * out * wj)
        gw_j.copy_(g_upd * out * wi)
        g_p = _gate_bwd(g_upd * wi * wj, saved, ln, g_ln)
        if g_p_out is not None:
            g_p_out.copy_(g_p)
        g_h = torch.cat([g_p[:, :64] @ w2[:64], g_p[:, 64:] @ w2[64:]], dim=1)
        g_pre.copy_(g_h * _dsilu(save_pre))

    # ---- K6
    def angle_update_fwd(self, pij, px, pa, ang, ang_atom, ang_i, ang_j, ln, ang_new, save_p):
        p = self._bond_pre(pij, px, pa, ang_atom, ang_i, ang_j)
        out, _ = _gate_fwd(p, ln)
        ang_new.copy_(out + ang)
        if save_p is not None:
            save_p.copy_(p)

    def angle_update_bwd(self, save_p, g_ang_in, ln, g_pre, g_ln=None):
        _, saved = _gate_fwd(save_p, ln)
        if g_ang_in is None:
            g_ang_in = torch.zeros_like(save_p[:, :64])
        g_pre.copy_(_gate_bwd(g_ang_in, saved, ln, g_ln))


    # ---- second-order kernels: tangent pass along rdot, and the reverse of (primal + tangent) --------
    def edge_tangent(self, rvec, dist, rhat, center, nbr, owner, u_atom, w_graph, ddist, drhat):
        """rdot_e = u[c] - u[n] + r_e . W[graph(c)];  ddist = rhat . rdot;  drhat = (rdot - rhat ddist)/d"""
        c, n = center.long(), nbr.long()
        rdot = u_atom[c] - u_atom[n] + torch.einsum("ei,eij->ej", rvec, w_graph.view(-1, 3, 3)[owner.long()[c]])
        dd = (rhat * rdot).sum(dim=1)
        ddist.copy_(dd)
        drhat.copy_((rdot - rhat * dd[:, None]) / dist[:, None])

    def bond_basis_tangent(self, dist, ddist, u2d, freq_ag, freq_bg, rc_ag, rc_bg, p, w3t, e0d, wagd, wbgd, tbasis):
        du, ddu = dist[u2d.long()], ddist[u2d.long()]
        _, dag = _rbf(du, freq_ag, rc_ag, p)
        _, dbg = _rbf(du, freq_bg, rc_bg, p)
        tag, tbg = dag * ddu[:, None], dbg * ddu[:, None]
        e0d.copy_(tag @ w3t[0]), wagd.copy_(tag @ w3t[1]), wbgd.copy_(tbg @ w3t[2])
        R = freq_ag.shape[0]
        tbasis.zero_()
        tbasis[:, :R] = tag
        tbasis[:, 32 : 32 + R] = tbg

    def bond_basis_bwd2(self, dist, ddist, u2d, freq_ag, freq_bg, rc_ag, rc_bg, p, w3, lam_e0, lam_wag, lam_wbg, g_freq):
        """g_freq += d/dfreq < lam, (dB/dd ddist) W >  (lam held fixed)"""
        du, ddu = dist[u2d.long()], ddist[u2d.long()]
        gb_ag = lam_e0 @ w3[0] + lam_wag @ w3[1]
        gb_bg = lam_wbg @ w3[2]
        g_freq[0] += (gb_ag * _rbf_d_dfreq(du, freq_ag, rc_ag, p) * ddu[:, None]).sum(dim=0).to(g_freq.dtype)
        g_freq[1] += (gb_bg * _rbf_d_dfreq(du, freq_bg, rc_bg, p) * ddu[:, None]).sum(dim=0).to(g_freq.dtype)

    def _theta_dot(self, rhat, drhat, ang_di, ang_dj):
        i, j = ang_di.long(), ang_dj.long()
        u = (rhat[i] * rhat[j]).sum(dim=1) * (1 - 1e-6)
        ud = ((drhat[i] * rhat[j]).sum(dim=1) + (rhat[i] * drhat[j]).sum(dim=1)) * (1 - 1e-6)
        return torch.acos(u), -ud / torch.sqrt(1 - u * u)

    def angle_basis_tangent(self, rhat, drhat, ang_di, ang_dj, freq, wt, a0d, tbasis):
        th, thd = self._theta_dot(rhat, drhat, ang_di, ang_dj)
        arg = th[:, None] * freq[None, :]
        fd = torch.cat([torch.zeros_like(th[:, None]), torch.cos(arg) * freq, -torch.sin(arg) * freq], dim=1)
        fd = fd * thd[:, None] / math.sqrt(math.pi)
        a0d.copy_(fd @ wt)
        tbasis.zero_()
        tbasis[:, : fd.shape[1]] = fd

    def angle_basis_bwd2(self, rhat, drhat, ang_di, ang_dj, freq, w, lam_a0, g_freq):
        th, thd = self._theta_dot(rhat, drhat, ang_di, ang_dj)
        arg = th[:, None] * freq[None, :]
        nf = freq.shape[0]
        gf = (lam_a0 @ w) / math.sqrt(math.pi)
        sn, cs = torch.sin(arg), torch.cos(arg)
        # d/dw [ w cos(w th) ] = cos - w th sin ;  d/dw [ -w sin(w th) ] = -(sin + w th cos)
        term = gf[:, 1 : 1 + nf] * (cs - arg * sn) - gf[:, 1 + nf :] * (sn + arg * cs)
        g_freq += (term * thd[:, None]).sum(dim=0).to(g_freq.dtype)

    @staticmethod
    def _w2(h, w2t):
        return torch.cat([h[:, :64] @ w2t[:, :64], h[:, 64:] @ w2t[:, 64:]], dim=1)

    @staticmethod
    def _w2_bwd(g, w2):
        return torch.cat([g[:, :64] @ w2[:64], g[:, 64:] @ w2[64:]], dim=1)

    def atom_conv_tan(self, pcn_d, pe_d, wag, wag_d, center, nbr, d2u, save_pre, save_p, w2t, ln, msg_d, pre_d, p_d):
        pre_d.copy_(self._atom_pre(pcn_d, pe_d, center, nbr, d2u))
        p_d.copy_(self._w2(_dsilu(save_pre) * pre_d, w2t))
        o, od, _ = _gate_tan(save_p, p_d, ln)
        u = d2u.long()
        msg_d.copy_(od * wag[u] + o * wag_d[u])

    def atom_conv_bwd2(self, save_pre, save_p, pre_d, p_d, g_p_lam, wag, wag_d, center, d2u, lam_agg, bar_agg, w2, ln,
                       bar_pre, bar_w, u_out, g_ln):
        c, u = center.long(), d2u.long()
        lam, bar, w, wd = lam_agg[c], bar_agg[c], wag[u], wag_d[u]
        o, od, cache = _gate_tan(save_p, p_d, ln)
        bar_w.copy_(bar * o + lam * od)
        uu = _gate_bwd2(bar * w + lam * wd, lam * w, cache, ln, g_ln)
        u_out.copy_(uu)
        bar_pre.copy_(_dsilu(save_pre) * self._w2_bwd(uu, w2) + _d2silu(save_pre) * pre_d * self._w2_bwd(g_p_lam, w2))

    def bond_conv_tan(self, pij_d, px_d, pa_d, wbg, wbg_d, ang_atom, ang_i, ang_j, save_pre, save_p, w2t, ln, upd_d,
                      pre_d, p_d):
        pre_d.copy_(self._bond_pre(pij_d, px_d, pa_d, ang_atom, ang_i, ang_j))
        p_d.copy_(self._w2(_dsilu(save_pre) * pre_d, w2t))
        o, od, _ = _gate_tan(save_p, p_d, ln)
        i, j = ang_i.long(), ang_j.long()
        upd_d.copy_(od * wbg[i] * wbg[j] + o * (wbg_d[i] * wbg[j] + wbg[i] * wbg_d[j]))

    def bond_conv_bwd2(self, save_pre, save_p, pre_d, p_d, g_p_lam, wbg, wbg_d, ang_i, ang_j, lam_agg, bar_agg, w2, ln,
                       bar_pre, bar_wi, bar_wj, u_out, g_ln):
        i, j = ang_i.long(), ang_j.long()
        lam, bar = lam_agg[i], bar_agg[i]
        wi, wj, wdi, wdj = wbg[i], wbg[j], wbg_d[i], wbg_d[j]
        o, od, cache = _gate_tan(save_p, p_d, ln)
        bar_wi.copy_(bar * o * wj + lam * (od * wj + o * wdj))
        bar_wj.copy_(bar * o * wi + lam * (od * wi + o * wdi))
        uu = _gate_bwd2(bar * wi * wj + lam * (wdi * wj + wi * wdj), lam * wi * wj, cache, ln, g_ln)
        u_out.copy_(uu)
        bar_pre.copy_(_dsilu(save_pre) * self._w2_bwd(uu, w2) + _d2silu(save_pre) * pre_d * self._w2_bwd(g_p_lam, w2))

    def angle_update_tan(self, pij_d, px_d, pa_d, ang_d, ang_atom, ang_i, ang_j, save_p, ln, ang_new_d, p_d):
        p_d.copy_(self._bond_pre(pij_d, px_d, pa_d, ang_atom, ang_i, ang_j))
        _, od, _ = _gate_tan(save_p, p_d, ln)
        ang_new_d.copy_(od + ang_d)

    def angle_update_bwd2(self, save_p, p_d, lam_ang, bar_ang, ln, bar_pre, g_ln):
        """lam_ang / bar_ang: lambda / R2 adjoint of ang_new (None = zero)."""
        _, _, cache = _gate_tan(save_p, p_d, ln)
        z = torch.zeros_like(save_p[:, :64])
        bar_pre.copy_(_gate_bwd2(z if bar_ang is None else bar_ang, z if lam_ang is None else lam_ang, cache, ln, g_ln))

    def readout_bwd2(self, x, xd, ln, mlp_wt, mlp_w, mlp_b, w_last, seed, bar_x, h_all, hd_all, gz_all, zbar_all,
                     g_h0, hbar0, xhat, xhatd):
        """Reverse of (readout, its tangent along xd) for the scalar  sum_i seed_i site_e_i + T_i,
        T_i = <d site_e_i/dx_i, xd_i>.  gz_all = lambda of the pre-activations (seed 1) = adjoint of the
        tangent pre-activations; zbar_all = adjoint of the primal pre-activations."""
        L = mlp_wt.shape[0]
        if ln is not None:
            h0, xh, rstd = _ln_fwd(x, ln[0], ln[1])
            xd_hat = _ln_tan(xd, xh, rstd)
            hd = ln[0] * xd_hat
            xhat.copy_(xh), xhatd.copy_(xd_hat)
        else:
            h0, hd = x, xd
        h, zs, zds = h0, [], []
        for l in range(L):
            h_all[l].copy_(h), hd_all[l].copy_(hd)
            z = h @ mlp_wt[l] + mlp_b[l]
            zd = hd @ mlp_wt[l]
            zs.append(z), zds.append(zd)
            h, hd = _silu(z), _dsilu(z) * zd
        h_all[L].copy_(h), hd_all[L].copy_(hd)
        hdbar = w_last[None, :].expand_as(h)  # adjoint of hd_L (T = hd_L . w_last)
        hbar = w_last[None, :] * seed[:, None]  # adjoint of h_L (energy-loss seed)
        for l in reversed(range(L)):
            zdbar = hdbar * _dsilu(zs[l])
            zbar = hdbar * _d2silu(zs[l]) * zds[l] + hbar * _dsilu(zs[l])
            gz_all[l].copy_(zdbar), zbar_all[l].copy_(zbar)
            hdbar, hbar = zdbar @ mlp_w[l], zbar @ mlp_w[l]
        g_h0.copy_(hdbar), hbar0.copy_(hbar)
        if ln is None:
            bar_x.copy_(hbar)
            return
        kk = hdbar * ln[0]
        v = hbar * ln[0] + rstd * (-kk * (xh * xd).mean(dim=1, keepdim=True) - (kk * xh).sum(dim=1, keepdim=True) * xd / 64)
        q = rstd * (v - v.mean(dim=1, keepdim=True) - xh * (v * xh).mean(dim=1, keepdim=True))
        bar_x.copy_(q - rstd * xh * (kk * xd_hat).sum(dim=1, keepdim=True) / 64)

    # ---- training-only kernels (reference trainer.py:398-411: loss.backward() + optimizer.step())
    def wgrad(self, x, g, out, colsum=None, x_rows=None, g_rows=None, x_silu=False, x2=None):
        """out[64][n] = act(x[x_rows])^T @ g[g_rows]; colsum[n] = sum_rows g[g_rows].  x, g, out may be
        column-slice views (unit column stride).  act: identity; x_silu -> silu(x); x2 given ->
        silu'(x) * x2 (the tangent of the hidden activations)."""
        xs = x if x_rows is None else x[x_rows.long()]
        gs = g if g_rows is None else g[g_rows.long()]
        if x2 is not None:
            xs = _dsilu(xs) * x2
        elif x_silu:
            xs = _silu(xs)
        out.copy_(xs.T @ gs)
        if colsum is not None:
            colsum.copy_(gs.sum(dim=0))

    def colsum(self, a, out, b=None, rowscale=None):
        """out[n] (fp64) = sum_rows a * b * rowscale[:, None]."""
        v = a.to(out.dtype)
        if b is not None:
            v = v * b.to(out.dtype)
        if rowscale is not None:
            v = v * rowscale.to(out.dtype)[:, None]
        out.copy_(v.sum(dim=0))

    def readout_bwd(self, x, ln, mlp_wt, mlp_w, mlp_b, w_last, seed, g_x, h_all, gz_all, g_h0, xhat):
        """Reverse of the readout MLP with a per-atom seed dL/d(site energy); also returns what the
        parameter gradients need: h_all[l] = input of linear l (l = L: input of the last linear),
        gz_all[l] = dL/d(pre-activation l), g_h0 = dL/d(LayerNorm output), xhat."""
        if ln is not None:
            h0, xh, rstd = _ln_fwd(x, ln[0], ln[1])
            xhat.copy_(xh)
        else:
            h0 = x
        acts, h = [], h0
        L = mlp_wt.shape[0]
        for l in range(L):
            h_all[l].copy_(h)
            zl = h @ mlp_wt[l] + mlp_b[l]
            acts.append(zl)
            h = _silu(zl)
        h_all[L].copy_(h)
        g = w_last[None, :] * seed[:, None]
        for l in reversed(range(L)):
            gz = g * _dsilu(acts[l])
            gz_all[l].copy_(gz)
            g = gz @ mlp_w[l]
        g_h0.copy_(g)
        if ln is not None:
            g = _ln_bwd(g, xh, rstd, ln[0])
        g_x.copy_(g)

    def magmom_bwd(self, x, w, b, g_m, g_x, g_lin):
        """m = |x.w + b|: g_lin[i] = dL/d(x_i.w + b) = sign * g_m[i]; g_x += g_lin (x) w."""
        g_lin.copy_(torch.sign(x @ w + b) * g_m)
        g_x.add_(g_lin[:, None] * w[None, :])

    def adam_step(self, p, g, m, v, lr, beta1, beta2, eps, weight_decay, step):
        """torch.optim.Adam (no amsgrad) on one flat buffer; L2 weight decay added to the gradient."""
        gg = g + weight_decay * p if weight_decay != 0 else g
        m.mul_(beta1).add_(gg, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
        bc1, bc2 = 1 - beta1**step, 1 - beta2**step
        p.sub_(lr / bc1 * m / ((v / bc2).sqrt() + eps))

    def loss_terms(self, pred, target, kind, delta, g_pred, sums):
        """One CombinedLoss term (trainer.py:779-869) over a flat vector with NaN-masked targets:
        sums (fp64 [3]) += [sum loss_i, sum |err_i|, count]; g_pred = d loss_i / d pred_i (0 where masked).
        kind 0 MSE, 1 MAE, 2 Huber(delta)."""
        valid = ~torch.isnan(target)
        err = torch.where(valid, pred - torch.nan_to_num(target), torch.zeros_like(pred))
        if kind == 0:
            li, gi = err**2, 2 * err
        elif kind == 1:
            li, gi = err.abs(), torch.sign(err)
        else:
            small = err.abs() <= delta
            li = torch.where(small, 0.5 * err**2, delta * (err.abs() - 0.5 * delta))
            gi = torch.where(small, err, delta * torch.sign(err))
        g_pred.copy_(torch.where(valid, gi, torch.zeros_like(gi)))
        sums += torch.stack([li.sum(), err.abs().sum(), valid.sum().to(li.dtype)]).to(sums.dtype)

    # ---- K7
    def readout(self, x, z, owner, ln, mlp_wt, mlp_w, mlp_b, w_last, b_last, atom_ref, site_e, h_out, e_graph, e_ref, g_x):
        if ln is not None:
            h0, xhat, rstd = _ln_fwd(x, ln[0], ln[1])
        else:
            h0 = x
        acts, h = [], h0
        for l in range(mlp_wt.shape[0]):
            zl = h @ mlp_wt[l] + mlp_b[l]
            acts.append(zl)
            h = _silu(zl)
        se = h @ w_last + b_last
        site_e.copy_(se)
        if h_out is not None:
            h_out.copy_(h0)
        e_graph.index_add_(0, owner.long(), se.to(e_graph.dtype))
        e_ref.index_add_(0, owner.long(), atom_ref[z.long() - 1].to(e_ref.dtype))
        if g_x is not None:
            g = w_last[None, :].expand_as(h).clone()
            for l in reversed(range(mlp_wt.shape[0])):
                g = (g * _dsilu(acts[l])) @ mlp_w[l]
            if ln is not None:
                g = _ln_bwd(g, xhat, rstd, ln[0])
            g_x.copy_(g)

    def magmom(self, x, w, b, m):
        m.copy_(torch.abs(x @ w + b))

    # ---- K1c
    def force_virial(self, rvec, dist, rhat, g_rhat, g_dist, d2u, u2d, center, nbr, owner, force, virial):
        gr = g_rhat.to(torch.float64)
        rh = rhat.to(torch.float64)
        d = dist.to(torch.float64)
        g = (gr - rh * (rh * gr).sum(dim=1, keepdim=True)) / d[:, None]
        u = d2u.long()
        is_rep = u2d.long()[u] == torch.arange(len(u))
        g = g + torch.where(is_rep, g_dist.to(torch.float64)[u], torch.zeros_like(d))[:, None] * rh
        force.index_add_(0, center.long(), -g)
        force.index_add_(0, nbr.long(), g)
        outer = rvec.to(torch.float64)[:, :, None] * g[:, None, :]
        virial.index_add_(0, owner.long()[center.long()], outer.reshape(-1, 9))
