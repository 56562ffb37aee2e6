"""Leave-beam-out refits (BASELINE.json configs[4]).

The reference's `--validate` fits a time window and draws a figure (validate.py:53-132); it has no refit study.
SURVEY row V defines the study this module runs: for every record r and every beam b, the SAME fit with beam b's
gates masked (value = error = NaN, exactly what the quality filter does to a bad gate, interpolate.py:655-657) --
so the reference fit (interpolate.py:511-574) on the masked input defines the expected result, and on the GPU it is the ordinary
batched fit over R x nbeams "virtual records".  Masking is exact (a masked gate is a zero-weight gate,
interpolate.py:516-520), records and beams are independent, so the virtual records shard and batch like real ones.

The held-out beam's prediction residual -- what a validation study is after -- comes with it: chi^2 of the refitted
model on the gates that were left out.
"""
from dataclasses import dataclass

import numpy as np

from . import _native, fit as _fit


@dataclass
class LeaveBeamOutResult:
    Coeffs: object        # (R, nbeams, N)
    chi_sq: object        # (R, nbeams)   chi^2 on the gates that were kept (interpolate.py:569)
    reg_params: object    # (R, nbeams, nreg)
    status: object        # (R, nbeams)
    rank: object          # (R, nbeams)
    heldout_chi_sq: object   # (R, nbeams) sum_j W_j (A C - b)_j^2 over the held-out beam's valid gates
    heldout_count: object    # (R, nbeams) number of those gates
    nsolve: int = 0


def beam_index(nbeams, ngates, keep):
    """Beam of every retained gate: gates are flattened beams-major (interpolate.py:635-642) and the NaN-altitude
    ones dropped (:660-664); `keep` is that boolean mask over nbeams * ngates."""
    return np.repeat(np.arange(nbeams), ngates)[np.asarray(keep, dtype=bool)]


def leave_beam_out(model, lat, lon, alt, value, error, beam_of_gate, reg_matrices=None, method='chi2',
                   ne_mode=_native.NE_FAST, device=None, virtual_batch=16384):
    """Fit every (record, beam) pair with that beam's gates masked.  value / error: (R, P) host or CUDA arrays;
    beam_of_gate: (P,) int.  Virtual records are processed `virtual_batch` at a time (whole records)."""
    import torch
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    f64 = lambda a: a.to(dev, torch.float64) if isinstance(a, torch.Tensor) else \
        torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    beam = torch.from_numpy(np.ascontiguousarray(beam_of_gate, dtype=np.int64)).to(dev)
    beams = torch.unique(beam)
    nb = int(beams.numel())
    R, P = value.shape
    N = model.nbasis
    nreg = 0 if not reg_matrices else len(reg_matrices)
    la, lo, al = f64(np.ravel(lat)), f64(np.ravel(lon)), f64(np.ravel(alt))
    A = model.basis_device(la, lo, al)
    held = (beam[None, :] == beams[:, None])                      # (nb, P): gates of the held-out beam
    per = max(1, virtual_batch // nb)
    nan = float('nan')
    out = LeaveBeamOutResult(*[None] * 7)
    C_all = torch.empty((R, nb, N), dtype=torch.float64, device=dev)
    chi_all = torch.empty((R, nb), dtype=torch.float64, device=dev)
    lam_all = torch.empty((R, nb, nreg), dtype=torch.float64, device=dev)
    st_all = torch.empty((R, nb), dtype=torch.int32, device=dev)
    rk_all = torch.empty((R, nb), dtype=torch.int32, device=dev)
    ho_all = torch.empty((R, nb), dtype=torch.float64, device=dev)
    hn_all = torch.empty((R, nb), dtype=torch.int32, device=dev)
    nsolve = 0
    for r0 in range(0, R, per):
        r1 = min(R, r0 + per)
        v, e = f64(value[r0:r1]), f64(error[r0:r1])
        # virtual records (r, b): beam b's gates -> NaN
        vv = v[:, None, :].expand(-1, nb, -1).masked_fill(held[None], nan).reshape(-1, P).contiguous()
        ee = e[:, None, :].expand(-1, nb, -1).masked_fill(held[None], nan).reshape(-1, P).contiguous()
        res = _fit.fit_records(model, lat, lon, alt, vv, ee, reg_matrices, method, ne_mode=ne_mode, device=dev,
                               rec_batch=vv.shape[0], to_host=False)
        nsolve += res.nsolve
        k = r1 - r0
        C_all[r0:r1] = res.Coeffs.view(k, nb, N)
        chi_all[r0:r1] = res.chi_sq.view(k, nb)
        if nreg:
            lam_all[r0:r1] = res.reg_params.view(k, nb, nreg)
        st_all[r0:r1] = res.status.view(k, nb)
        rk_all[r0:r1] = res.rank.view(k, nb)
        # prediction on the held-out gates: (k nb) x P densities, only the held-out beam's valid gates count
        pred = (res.Coeffs @ A.t()).view(k, nb, P)
        ok = held[None] & torch.isfinite(v)[:, None, :]
        w = torch.where(ok, e[:, None, :].expand(-1, nb, -1) ** -2, torch.zeros((), dtype=torch.float64, device=dev))
        d = torch.where(ok, pred - v[:, None, :], torch.zeros((), dtype=torch.float64, device=dev))
        ho_all[r0:r1] = (d * d * w).sum(dim=2)
        hn_all[r0:r1] = ok.sum(dim=2).to(torch.int32)
        del vv, ee, res, pred, w, d
    host = lambda t: t.cpu().numpy()
    return LeaveBeamOutResult(host(C_all), host(chi_all), host(lam_all), host(st_all), host(rk_all), host(ho_all),
                              host(hn_all), nsolve)
