"""Drop-in for the DSVI line of the reference (code/nmgp_dsvi.py): ``NMGP``, ``inference``,
``predict_Y``, ``vec2list``, ``pre_intialization`` with the reference's signatures.

Differences a user can see:
* parameters and data live on the CUDA device; the arithmetic runs in the sm_100a kernels of
  ``libnmgp_b200.so`` (no CPU fallback -- a CPU model raises on ``forward``);
* ``forward`` returns a tensor whose ``backward`` delivers the hand-written gradient (no autograd graph
  of thousands of nodes); ``n_mc`` (default 1, the reference's value) averages that many reparameterised
  draws in one call;
* ``noise="reference"`` (default) consumes the global CPU generator exactly as the reference does
  (float32 draws cast to float64, order z_v, z_ell, z_ij for i>=j: code/utils.py:123,226,234), so seeded
  runs reproduce the reference's loss trace; ``noise="device"`` draws only the B*(D+1)/2 normals the
  estimator actually uses, on the GPU.
"""
from __future__ import annotations

import time
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch.nn import Parameter
from torch.utils.data import Dataset

from . import _ops as ops
from . import dsvi_step as _step

TensorType = torch.DoubleTensor           # code/nmgp_dsvi.py:18
F64 = torch.float64


def default_device() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


class trainData(Dataset):
    """code/nmgp_dsvi.py:86-96."""

    def __init__(self, X_data, Y_data, I):
        self.X_data, self.Y_data, self.I = X_data, Y_data, I

    def __getitem__(self, index):
        return self.X_data[index], self.Y_data[index], self.I[index]

    def __len__(self):
        return len(self.X_data)


def _rows_from_lists(inputs_list, outputs_list, D, index=None, subjects=False):
    """Concatenate per-output lists into (x, y, I) sorted by output id (stable); also returns the
    permutation applied so callers can restore the caller's row order.  ``subjects``: outputs_list[d] is [S, T_d]
    (S subjects observed on the same inputs) and y comes back as [S, B]."""
    ids = list(range(D)) if index is None else list(index)
    sizes = [int(x.shape[0]) for x in inputs_list]
    I = np.repeat(np.asarray(ids[:len(sizes)], dtype=np.int64), sizes)
    x = torch.cat([t.reshape(-1) for t in inputs_list]) if len(inputs_list) else torch.empty(0, dtype=F64)
    if outputs_list is None:
        y = None
    elif subjects:
        y = torch.cat([t.reshape(t.shape[0], -1) for t in outputs_list], dim=1)
    else:
        y = torch.cat([t.reshape(-1) for t in outputs_list])
    if I.size and np.any(np.diff(I) < 0):
        perm = np.argsort(I, kind="stable")
    else:
        perm = None
    return x, y, I, perm


class _DSVILoss(torch.autograd.Function):
    """Loss and gradient are produced together by the fused step; backward only rescales."""

    @staticmethod
    def forward(ctx, model, x, y, I, N, z_v, z_ell, z_L, step_kw, *params):
        p = dict(zip(_step.PARAM_NAMES, params))
        aux = None
        if step_kw.get("defer_pd_check"):
            aux = step_kw.get("aux")
            if aux is None:
                aux = {}
                step_kw = dict(step_kw, aux=aux)
        loss, grads = _step.dsvi_step(p, model.Z.reshape(-1), x, y, I, N, z_v, z_ell, z_L, **step_kw)
        model._last_pd_info = aux.get("pd_info") if aux is not None else None
        ctx.grads = [grads[k].reshape(p[k].shape) for k in _step.PARAM_NAMES]
        return loss

    @staticmethod
    def backward(ctx, gout):
        out = [None] * 9
        for g in ctx.grads:
            out.append(g * gout)
        return tuple(out)


class NMGP(torch.nn.Module):
    """code/nmgp_dsvi.py:99-155 (constructor, parameters, init rules, seed)."""

    def __init__(self, number_observations, dim_outputs, Z, minibatch_size=None, mu_v=None, mu_W=None, mu_U=None,
                 sqrt_v=None, sqrt_W=None, sqrt_U=None, seed=22, device=None, noise="reference", exact_kl=False):
        super().__init__()
        dev = default_device() if device is None else torch.device(device)
        self.Z = torch.as_tensor(Z, dtype=F64).to(dev)
        self.M = self.Z.shape[0]
        self.N = number_observations
        self.D = dim_outputs
        self.batch_size = minibatch_size
        self.noise = noise
        self.noise_seed = seed            # key of the counter-based device noise ("device" mode)
        self._noise_step = 0
        # exact_kl=True: mathematically correct KL terms instead of the reference's (quirk q10) -- explicit opt-in
        self.step_options = {"exact_kl": True} if exact_kl else {}
        D, M = self.D, self.M

        torch.random.manual_seed(seed)            # same draw order as the reference (CPU generator)
        sqrt_scale = 0.1

        def given(a):
            return torch.from_numpy(np.asarray(a)).type(TensorType)
        mu_W_t = 0.1 * torch.randn(D, M).type(TensorType) if mu_W is None else given(mu_W)
        sqrt_W_t = sqrt_scale * torch.randn(D, M, M).type(TensorType) if sqrt_W is None else given(sqrt_W)
        mu_v_t = -4 * torch.ones(M).type(TensorType) if mu_v is None else given(mu_v)
        sqrt_v_t = sqrt_scale * torch.randn(M, M).type(TensorType) if sqrt_v is None else given(sqrt_v)
        mu_U_t = 0.1 * torch.randn(D, D, M).type(TensorType) if mu_U is None else given(mu_U)
        sqrt_U_t = sqrt_scale * torch.randn(D, D, M, M).type(TensorType) if sqrt_U is None else given(sqrt_U)
        # registration order == the reference's, so state_dict()/optimizer param order match (quirk q8:
        # the j>i blocks of mu_U/sqrt_U are parameters with identically zero gradient)
        self.mu_W = Parameter(mu_W_t.to(dev))
        self.sqrt_W = Parameter(sqrt_W_t.to(dev))
        self.mu_v = Parameter(mu_v_t.to(dev))
        self.sqrt_v = Parameter(sqrt_v_t.to(dev))
        self.mu_U = Parameter(mu_U_t.to(dev))
        self.sqrt_U = Parameter(sqrt_U_t.to(dev))
        self.sigma2_g = 1
        mk = lambda v: Parameter(torch.tensor(v, dtype=F64, device=dev))
        self.sigma2_tildeell_log = mk(0.)
        self.length_scales_tildeell_log = mk(-4.)
        self.sigma2_L0_log = mk(0.)
        self.length_scales_L0_log = mk(-4.)
        self.sigma2_L1_log = mk(0.)
        self.length_scales_L1_log = mk(-4.)
        self.sigma2_err_log = mk(-2.)

    # -- helpers ---------------------------------------------------------------------------------
    @property
    def device(self):
        return self.mu_W.device

    def _param_list(self):
        return [getattr(self, k) for k in _step.PARAM_NAMES]

    def _reference_noise(self, B, I_sorted, perm, n_mc):
        """Draw in the reference's order on the CPU generator.  The reference draws z_ij for the rows in
        the caller's order; ``perm`` maps sorted rows back to that order."""
        D, Q = self.D, self.M
        zv = torch.empty(n_mc, Q, dtype=F64); zell = torch.empty(n_mc, B, dtype=F64)
        zL = torch.zeros(n_mc, B, D, dtype=F64)
        It = torch.from_numpy(I_sorted)
        sel = [torch.nonzero(It == i).reshape(-1) for i in range(D)]
        pidx = None if perm is None else torch.from_numpy(perm)
        for s in range(n_mc):
            zv[s] = torch.randn(Q).type(TensorType)
            ze = torch.randn(B).type(TensorType)
            zell[s] = ze if pidx is None else ze[pidx]
            for i in range(D):
                for j in range(i + 1):
                    z = torch.randn(B).type(TensorType)
                    if sel[i].numel():
                        zs = z if pidx is None else z[pidx]
                        zL[s, sel[i], j] = zs[sel[i]]
        return zv, zell, zL

    def _device_noise(self, B, n_mc, row_gid=None, sample_offset=0):
        """Counter-based device noise (csrc/philox.cuh): float32 normals widened to float64 like the reference's
        (quirk q2), a pure function of (noise_seed, step, sample, global row id, column) -- identical however the rows
        are sharded over ranks.  z_v and z_ell are materialised (small); the B*D coefficient draws are generated inside
        the sampling kernels and never stored (returned as None plus the key)."""
        dev = self.device
        seed = int(self.noise_seed)
        sid = None if not sample_offset else torch.arange(sample_offset, sample_offset + n_mc, dtype=torch.int64, device=dev)
        ctr = getattr(self, "_noise_step_dev", None)
        if ctr is not None:
            # CUDA-graph mode: the step index lives on the device and is bumped by the (captured) step itself
            base, step_dev = 0, ctr
        else:
            base, step_dev = self._noise_step << 8, None
            self._noise_step += 1
        zv = ops.noise_fill(1, n_mc, self.M, seed, base | 0, 0, sid, dev, step_dev=step_dev)[0]
        zell = ops.noise_fill(n_mc, B, 1, seed, base | 1, int(sample_offset), row_gid, dev, step_dev=step_dev).reshape(n_mc, B)
        return zv, zell, None, (seed, base | 2, step_dev)

    def use_device_step_counter(self, enable=True):
        """Keep the noise step counter on the device (needed to replay a step from a CUDA graph: kernel arguments are
        frozen at capture, so the step index must be data).  The caller bumps it with ``advance_noise_step()`` inside
        the captured region.  Draws are identical to the host-counter mode."""
        if enable:
            if getattr(self, "_noise_step_dev", None) is None:       # idempotent: a captured graph holds this buffer
                self._noise_step_dev = torch.full((1,), int(self._noise_step), dtype=torch.int64, device=self.device)
        else:
            if getattr(self, "_noise_step_dev", None) is not None:
                self._noise_step = int(self._noise_step_dev.item())
            self._noise_step_dev = None

    def advance_noise_step(self):
        if getattr(self, "_noise_step_dev", None) is not None:
            self._noise_step_dev.add_(1)

    def forward_rows(self, x, y, I, n_mc=1, explicit_noise=None, row_gid=None):
        """Same as forward() for rows already on the device: x, y float64 [B], I int32 [B] sorted by output;
        ``row_gid`` (int64 [B]) are the rows' global ids when the minibatch is sharded over ranks."""
        kw = dict(self.step_options)
        if explicit_noise is not None:
            zv, zell, zL = explicit_noise
        else:
            zv, zell, zL, key = self._device_noise(x.shape[0], n_mc, row_gid, int(kw.get("sample_offset", 0)))
            kw.update(noise_key=key, row_gid=row_gid)
        return _DSVILoss.apply(self, x, y, I, self.N, zv, zell, zL, kw, *self._param_list())

    # -- the hot path -----------------------------------------------------------------------------
    def forward(self, inputs_list, outputs_list, index=None, verbose=False, n_mc=1, noise=None, explicit_noise=None,
                row_gid=None, subjects=False):
        """-SELBO of one minibatch (code/nmgp_dsvi.py:157-301).  ``subjects=True`` (not in the reference):
        outputs_list[d] is [S, T_d] -- S subjects observed on the same inputs; the result is the mean over the subjects
        of the reference's one-draw forward on each subject's targets (n_mc must equal S)."""
        t1 = time.time() if verbose else None
        x, y, I, perm = _rows_from_lists(inputs_list, outputs_list, self.D, index, subjects)
        B = x.shape[0]
        if subjects and y.shape[0] != n_mc:
            raise ValueError("subjects=True needs n_mc == number of subjects (%d != %d)" % (n_mc, y.shape[0]))
        if perm is not None:
            pt = torch.from_numpy(perm)
            x, y, I = x[pt.to(x.device)], (y[:, pt.to(y.device)] if subjects else y[pt.to(y.device)]), I[perm]
        noise = noise or self.noise
        dev = self.device
        kw = dict(self.step_options)
        if explicit_noise is not None:
            zv, zell, zL = explicit_noise
        elif noise == "reference":
            zv, zell, zL = self._reference_noise(B, I, perm, n_mc)
        elif noise == "device":
            zv, zell, zL, key = self._device_noise(B, n_mc, row_gid, int(kw.get("sample_offset", 0)))
            kw.update(noise_key=key, row_gid=row_gid)
        else:
            raise ValueError("noise must be 'reference' or 'device'")
        up = lambda t: None if t is None else t.to(dev, dtype=F64, non_blocking=True).contiguous()
        Id = torch.from_numpy(I.astype(np.int32)).to(dev, non_blocking=True)
        loss = _DSVILoss.apply(self, up(x), up(y), Id, self.N, up(zv), up(zell), up(zL), kw, *self._param_list())
        if verbose:
            torch.cuda.synchronize()
            print("forward+gradient (fused) costs {}s".format(time.time() - t1))
        return loss

    # -- deterministic posterior mean -----------------------------------------------------------------
    def predict_Y(self, inputs_list, index=None):
        """code/nmgp_dsvi.py:666-722: E[L](x) E[g](x) at each row's own output, no sampling."""
        from .predict import posterior_mean
        x, _, I, perm = _rows_from_lists(inputs_list, None, self.D, index)
        dev = self.device
        if perm is not None:
            xs, Is = x[torch.from_numpy(perm)], I[perm]
        else:
            xs, Is = x, I
        p = {k: getattr(self, k).detach() for k in _step.PARAM_NAMES}
        out = posterior_mean(p, self.Z.reshape(-1), xs.to(dev, dtype=F64).contiguous(),
                             torch.from_numpy(Is.astype(np.int32)).to(dev))
        if perm is not None:
            inv = torch.empty_like(out)
            inv[torch.from_numpy(perm).to(dev)] = out
            out = inv
        return out

    def sample_Y(self, inputs_list, index=None, n_sample=1000, **kw):
        """code/nmgp_dsvi.py:406-491 -> (sampled_Ys [S,B], sampled_Ls [S,B,D], sampled_Gs [S,D,B], tilde_ells [S,B])."""
        from .predict import sample_Y as _sy
        return _sy(self, inputs_list, index=index, n_sample=n_sample, **kw)

    sample_Y_gpu = sample_Y          # code/nmgp_dsvi.py:582-664: the reference's device variant of the same sampler

    def sample_FY(self, inputs, n_sample=1000, **kw):
        """code/nmgp_dsvi.py:493-580 -> (tilde_ells [S,B], Ys [S,B,D], corrs [S,B,D,D])."""
        from .predict import sample_FY as _sf
        return _sf(self, inputs, n_sample=n_sample, **kw)

    def compute_ELBO(self, inputs_list, outputs_list, index=None, n_sample=1000, verbose=False, **kw):
        """code/nmgp_dsvi.py:303-404 (quirk q5 reproduced); extra keywords: noise, chunk."""
        from .predict import mc_elbo
        return mc_elbo(self, inputs_list, outputs_list, index=index, n_sample=n_sample, verbose=verbose, **kw)


# ------------------------------------------------------------------------------------------------------
def pre_intialization(M, D, factor=1e-2):
    """code/nmgp_dsvi.py:737-742 (name kept, including its spelling)."""
    mu_W = np.zeros([D, M])
    sqrt_v = np.eye(M) * factor
    sqrt_W = np.stack([np.eye(M) for _ in range(D)]) * factor
    sqrt_U = np.stack([np.stack([np.eye(M) for _ in range(D)]) for _ in range(D)]) * factor
    return mu_W, sqrt_v, sqrt_W, sqrt_U


def vec2list(X, Y, I, dim, device=None):
    """code/nmgp_dsvi.py:745-755: regroup flat rows by output id."""
    X_list, Y_list = [], []
    for m in range(dim):
        sel = I == m
        xs, ys = X[sel], Y[sel]
        if device is not None:
            xs, ys = xs.to(device), ys.to(device)
        X_list.append(xs); Y_list.append(ys)
    return X_list, Y_list


class DeviceMinibatches:
    """Shuffled minibatches gathered ON THE DEVICE (replaces the reference's CPU ``DataLoader(trainData, shuffle=True)`` +
    ``vec2list`` regrouping, code/nmgp_dsvi.py:816-837): the flattened training rows are uploaded once;
    per epoch only a permutation is drawn and per batch only an index vector crosses PCIe, the rows are gathered by
    ``index_select`` on the GPU already grouped by output (the order ``vec2list`` produces).

    ``order="reference"``: the permutation is drawn exactly as the reference's loader draws it -- from the global CPU
    generator, one base-seed draw for the loader iterator and one seed for the sampler's private generator -- so a
    seeded run sees the reference's batches and, with ``noise="reference"``, reproduces its loss trace (quirk q9).
    ``order="device"``: ``torch.randperm`` on the GPU from a private device generator; nothing is drawn on the host."""

    def __init__(self, X, Y, I, dim_outputs, batch_size, device, order="reference", seed=0):
        self.n = int(X.shape[0])
        self.bs = int(batch_size)
        self.D = int(dim_outputs)
        self.dev = torch.device(device)
        self.order = order
        self.I_host = np.asarray(I).reshape(-1).astype(np.int64)
        self.X = torch.as_tensor(X, dtype=F64).reshape(-1).to(self.dev)
        self.Y = torch.as_tensor(Y, dtype=F64).reshape(-1).to(self.dev)
        self.I = torch.from_numpy(self.I_host.astype(np.int32)).to(self.dev)
        if order == "device":
            self.gen = torch.Generator(device=self.dev)
            self.gen.manual_seed(int(seed))
        elif order != "reference":
            raise ValueError("order must be 'reference' or 'device'")

    def __len__(self):
        return (self.n + self.bs - 1) // self.bs

    @staticmethod
    def loader_permutation(n):
        """The permutation ``DataLoader(dataset, shuffle=True)`` would use for its next epoch, consuming the global CPU
        generator in the same way (checked against the real DataLoader in tests/test_api_hostlogic_cpu.py)."""
        torch.empty((), dtype=torch.int64).random_()                       # base seed of the loader's iterator
        seed = int(torch.empty((), dtype=torch.int64).random_().item())     # seed of the RandomSampler's generator
        g = torch.Generator()
        g.manual_seed(seed)
        return torch.randperm(n, generator=g)

    def epoch(self):
        """Yields (x, y, I_sorted [int32, device], I_sorted_host or None) per batch, rows grouped by output (stable)."""
        if self.order == "reference":
            perm = self.loader_permutation(self.n).numpy()
            for k in range(0, self.n, self.bs):
                idx = perm[k:k + self.bs]
                Ib = self.I_host[idx]
                grp = np.argsort(Ib, kind="stable")
                rows = torch.from_numpy(idx[grp]).to(self.dev, non_blocking=True)
                yield self.X.index_select(0, rows), self.Y.index_select(0, rows), self.I.index_select(0, rows), Ib[grp]
        else:
            perm = torch.randperm(self.n, device=self.dev, generator=self.gen)
            for k in range(0, self.n, self.bs):
                idx = perm[k:k + self.bs]
                Ib = self.I.index_select(0, idx)
                grp = torch.sort(Ib, stable=True)[1]
                rows = idx.index_select(0, grp)
                yield self.X.index_select(0, rows), self.Y.index_select(0, rows), Ib.index_select(0, grp), None


def _apply_hyperpars(model, hyperpars, fix_hyperpars):
    """code/nmgp_dsvi.py:779-814 with its quirks: q3 ('sigma2_L1_log' is written into sigma2_L0_log, :784-785) and q4
    (hyperpars=None with fix_hyperpars=True raises TypeError at the membership test, :809)."""
    targets = {"sigma2_tildeell_log": "sigma2_tildeell_log", "sigma2_L0_log": "sigma2_L0_log",
               "sigma2_L1_log": "sigma2_L0_log", "sigma2_err_log": "sigma2_err_log"}
    if hyperpars is not None:
        for key, dest in targets.items():                                   # dict order = the reference's statement order
            if key in hyperpars:
                getattr(model, dest).data.fill_(hyperpars[key])
    if fix_hyperpars:
        for name in ("length_scales_tildeell_log", "length_scales_L0_log", "length_scales_L1_log"):
            getattr(model, name).requires_grad = False
            if name in hyperpars:                                           # TypeError when hyperpars is None (q4)
                getattr(model, name).data.fill_(hyperpars[name])


def inference(X_train_list, Y_train_list, z, batch_size, dim_outputs, hyperpars=None, fix_hyperpars=True, mu_v=None,
              mu_W=None, mu_U=None, sqrt_v=None, sqrt_W=None, sqrt_U=None, lr=0.01, itnum=1000,
              do_stop_criterion=False, seed=22, verbose=False, PATH="model.pt", continuous_training=False,
              show_ELBO=True, save_model=False, X_test_list=None, Y_test_list=None, n_mc=1, noise="reference",
              device=None, batch_order=None):
    """Drop-in for code/nmgp_dsvi.py:758-909: same arguments and returns (+ n_mc, noise, device, batch_order).
    The training rows live on the GPU and the shuffled minibatches are gathered there (``DeviceMinibatches``);
    ``batch_order`` defaults to "reference" when ``noise == "reference"`` (seeded runs then reproduce the reference's
    batches and loss trace) and to "device" otherwise."""
    dev = default_device() if device is None else torch.device(device)
    X_all = np.concatenate(X_train_list).reshape(-1)
    Y_all = np.concatenate(Y_train_list).reshape(-1)
    I_all = np.concatenate([np.full(len(Y_train_list[d]), d, dtype=np.int64) for d in range(dim_outputs)])
    n_train = int(Y_all.shape[0])
    Z = torch.from_numpy(np.asarray(z)).type(TensorType).unsqueeze(1)

    model = NMGP(number_observations=n_train, dim_outputs=dim_outputs, Z=Z, minibatch_size=batch_size,
                 mu_v=mu_v, mu_W=mu_W, mu_U=mu_U, sqrt_v=sqrt_v, sqrt_W=sqrt_W, sqrt_U=sqrt_U, seed=seed, device=dev,
                 noise=noise)
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    if continuous_training:                        # the reference overrides, then restores the checkpoint, then freezes
        _apply_hyperpars(model, hyperpars, False)
        checkpoint = torch.load(PATH, map_location=dev, weights_only=False)
        model.load_state_dict(checkpoint["model_state_dict"])
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        if fix_hyperpars:
            only_len = None if hyperpars is None else {k: v for k, v in hyperpars.items() if k.startswith("length_scales")}
            _apply_hyperpars(model, only_len, True)
    else:
        _apply_hyperpars(model, hyperpars, fix_hyperpars)

    batches = DeviceMinibatches(X_all, Y_all, I_all, dim_outputs, batch_size, dev,
                                order=batch_order or ("reference" if noise == "reference" else "device"), seed=seed)
    full_lists = None
    loss_list, time_list, rmse_test_list = [], [], []
    Y_test_vec = np.concatenate(Y_test_list) if X_test_list is not None else None
    steps_per_epoch = len(batches)
    t_start = time.time()
    epoch, loss = -1, None
    with torch.cuda.device(dev) if dev.type == "cuda" else _nullcontext():
        for epoch in range(itnum):
            for step, (xb, yb, Ib, Ib_host) in enumerate(batches.epoch(), start=1):
                optimizer.zero_grad()
                if noise == "reference":
                    explicit = tuple(t.to(dev) for t in model._reference_noise(int(xb.shape[0]), Ib_host, None, n_mc))
                    loss = model.forward_rows(xb, yb, Ib, n_mc=n_mc, explicit_noise=explicit)
                else:
                    loss = model.forward_rows(xb, yb, Ib, n_mc=n_mc)
                loss.backward()
                optimizer.step()
                loss_value = loss.detach().data.cpu().numpy()
                loss_list.append(loss_value)
                time_list.append(time.time() - t_start)
                if X_test_list is not None:
                    est = predict_Y(model, X_test_list)
                    rmse_test_list.append(np.sqrt(np.mean((est[:, None] - Y_test_vec) ** 2)))
                if verbose:
                    print("epoch: {}/{}, batch: {}/{}, loss: {}".format(epoch, itnum, step, n_train / batch_size, loss_value))
            if do_stop_criterion and epoch % 5 == 4 and epoch > 5:
                trace = np.array(loss_list)
                if trace[-steps_per_epoch:].sum() > trace[-6 * steps_per_epoch:-5 * steps_per_epoch].sum():
                    print("Stop criteria is satisfied.")
                    break
            if show_ELBO and epoch % 100 == 99:
                full_lists = full_lists or _lists_by_output(X_all, Y_all, I_all, dim_outputs)
                print("epoch: {}, ELBO: {}".format(epoch + 1, model.compute_ELBO(*full_lists).detach()))
    print("training takes {}s".format(time.time() - t_start))

    if save_model:
        torch.save({'epoch': epoch, 'model_state_dict': model.state_dict(),
                    'optimizer_state_dict': optimizer.state_dict(), 'loss': loss}, PATH)
    if show_ELBO:
        full_lists = full_lists or _lists_by_output(X_all, Y_all, I_all, dim_outputs)
        print("epoch: {}, ELBO: {}".format(epoch + 1, model.compute_ELBO(*full_lists).detach()))
    if X_test_list is not None:
        return model, loss_list, rmse_test_list, time_list
    return model, loss_list, time_list


def _lists_by_output(X, Y, I, D):
    Xt, Yt = torch.from_numpy(X).type(TensorType), torch.from_numpy(Y).type(TensorType)
    return ([Xt[torch.from_numpy(I == d)].view(-1, 1) for d in range(D)],
            [Yt[torch.from_numpy(I == d)].view(-1, 1) for d in range(D)])


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def sample_Y(model, X_list, n_sample=1000):
    """code/nmgp_dsvi.py:912-918."""
    X_list = [torch.from_numpy(np.asarray(x)).type(TensorType) for x in X_list]
    Ys, Ls, Gs, ells = model.sample_Y(X_list, n_sample=n_sample)
    return Ys.data.cpu().numpy(), Ls.data.cpu().numpy(), Gs.data.cpu().numpy(), ells.data.cpu().numpy()


def sample_FY(model, x, n_sample=1000):
    """code/nmgp_dsvi.py:921-924 (the reference unpacks the model's (tilde_ells, Ys, corrs) into differently named
    variables; the positional order of the returned arrays is kept)."""
    x = torch.from_numpy(np.asarray(x)).type(TensorType)
    a, b, c = model.sample_FY(x, n_sample=n_sample)
    return a.data.cpu().numpy(), b.data.cpu().numpy(), c.data.cpu().numpy()


def predict_Y(model, X_list):
    """code/nmgp_dsvi.py:927-930."""
    X_list = [torch.from_numpy(np.asarray(x)).type(TensorType) for x in X_list]
    return model.predict_Y(X_list).data.cpu().numpy()
