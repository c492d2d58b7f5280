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
from torch.utils.data import DataLoader, Dataset

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


def _rows_from_lists(inputs_list, outputs_list, D, index=None):
    """Concatenate per-output lists into (x, y, I) sorted by output id (stable); also returns the
    permutation applied so callers can restore the caller's row order."""
    ids = list(range(D)) if index is None else list(index)
    sizes = [int(x.shape[0]) for x in inputs_list]
    I = np.repeat(np.asarray(ids[:len(sizes)], dtype=np.int64), sizes)
    x = torch.cat([t.reshape(-1) for t in inputs_list]) if len(inputs_list) else torch.empty(0, dtype=F64)
    y = None if outputs_list is None else torch.cat([t.reshape(-1) for t in outputs_list])
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
        step = self._noise_step
        self._noise_step += 1
        seed = int(self.noise_seed)
        sid = None if not sample_offset else torch.arange(sample_offset, sample_offset + n_mc, dtype=torch.int64, device=dev)
        zv = ops.noise_fill(1, n_mc, self.M, seed, (step << 8) | 0, 0, sid, dev)[0]
        zell = ops.noise_fill(n_mc, B, 1, seed, (step << 8) | 1, int(sample_offset), row_gid, dev).reshape(n_mc, B)
        return zv, zell, None, (seed, (step << 8) | 2)

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
                row_gid=None):
        """-SELBO of one minibatch (code/nmgp_dsvi.py:157-301)."""
        t1 = time.time() if verbose else None
        x, y, I, perm = _rows_from_lists(inputs_list, outputs_list, self.D, index)
        B = x.shape[0]
        if perm is not None:
            x, y, I = x[torch.from_numpy(perm).to(x.device)], y[torch.from_numpy(perm).to(y.device)], I[perm]
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


def inference(X_train_list, Y_train_list, z, batch_size, dim_outputs, hyperpars=None, fix_hyperpars=True, mu_v=None,
              mu_W=None, mu_U=None, sqrt_v=None, sqrt_W=None, sqrt_U=None, lr=0.01, itnum=1000,
              do_stop_criterion=False, seed=22, verbose=False, PATH="model.pt", continuous_training=False,
              show_ELBO=True, save_model=False, X_test_list=None, Y_test_list=None, n_mc=1, noise="reference",
              device=None):
    """code/nmgp_dsvi.py:758-909, same arguments/returns (+ n_mc, noise, device).  Quirks kept: q3 (the
    'sigma2_L1_log' override lands in sigma2_L0_log, :784-785) and q4 (hyperpars=None with fix_hyperpars=True
    raises TypeError, :809)."""
    X_train_vec = np.concatenate(X_train_list)
    Y_train_vec = np.concatenate(Y_train_list)
    train_index = np.concatenate([np.ones_like(Y_train_list[i]) * i for i in range(dim_outputs)]).astype(int)
    X = torch.from_numpy(X_train_vec).type(TensorType)
    Y = torch.from_numpy(Y_train_vec).type(TensorType)
    I = torch.from_numpy(train_index).type(TensorType)
    dev = default_device() if device is None else torch.device(device)
    Z = torch.from_numpy(np.asarray(z)).type(TensorType).unsqueeze(1)
    X_list, Y_list = vec2list(X, Y, I, dim=dim_outputs)

    model = NMGP(number_observations=Y_train_vec.shape[0], dim_outputs=dim_outputs, Z=Z, minibatch_size=batch_size,
                 mu_v=mu_v, mu_W=mu_W, mu_U=mu_U, sqrt_v=sqrt_v, sqrt_W=sqrt_W, sqrt_U=sqrt_U, seed=seed, device=dev,
                 noise=noise)
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)

    if hyperpars is not None:
        if "sigma2_tildeell_log" in hyperpars:
            model.sigma2_tildeell_log.data.fill_(hyperpars['sigma2_tildeell_log'])
        if "sigma2_L0_log" in hyperpars:
            model.sigma2_L0_log.data.fill_(hyperpars['sigma2_L0_log'])
        if "sigma2_L1_log" in hyperpars:
            model.sigma2_L0_log.data.fill_(hyperpars['sigma2_L1_log'])       # quirk q3, kept
        if "sigma2_err_log" in hyperpars:
            model.sigma2_err_log.data.fill_(hyperpars['sigma2_err_log'])

    def freeze_lengthscales():
        for name in ("length_scales_tildeell_log", "length_scales_L0_log", "length_scales_L1_log"):
            getattr(model, name).requires_grad = False
            if name in hyperpars:                                           # TypeError when hyperpars is None (q4)
                getattr(model, name).data.fill_(hyperpars[name])

    if continuous_training:
        checkpoint = torch.load(PATH, map_location=dev, weights_only=False)
        model.load_state_dict(checkpoint["model_state_dict"])
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        if fix_hyperpars:
            freeze_lengthscales()
    elif fix_hyperpars:
        freeze_lengthscales()

    train_loader = DataLoader(trainData(X, Y, I), batch_size=batch_size, shuffle=True)
    loss_list, time_list = [], []
    if X_test_list is not None:
        rmse_test_list = []
        Y_test_vec = np.concatenate(Y_test_list)
    ts = time.time()
    epoch = -1
    loss = None
    for epoch in range(itnum):
        batch = 0
        for X_batch, Y_batch, I_batch in train_loader:
            batch += 1
            optimizer.zero_grad()
            X_batch_list, Y_batch_list = vec2list(X_batch, Y_batch, I_batch, dim=dim_outputs)
            loss = model(X_batch_list, Y_batch_list, verbose=verbose, n_mc=n_mc)
            loss.backward()
            optimizer.step()
            loss_value = loss.detach().data.cpu().numpy()
            loss_list.append(loss_value)
            time_list.append(time.time() - ts)
            if X_test_list is not None:
                est_Y_test = predict_Y(model, X_test_list)
                rmse_test_list.append(np.sqrt(np.mean((est_Y_test[:, None] - Y_test_vec) ** 2)))
            if verbose:
                print("epoch: {}/{}, batch: {}/{}, loss: {}".format(epoch, itnum, batch,
                                                                    X_train_vec.shape[0] / batch_size, loss_value))
        if do_stop_criterion:
            if epoch % 5 == 4 and epoch > 5:
                loss_array = np.array(loss_list)
                if np.sum(loss_array[-batch:]) > np.sum(loss_array[-batch * 6:-batch * 5]):
                    print("Stop criteria is satisfied.")
                    break
        if epoch % 100 == 99 and show_ELBO:
            elbo = model.compute_ELBO(X_list, Y_list)
            print("epoch: {}, ELBO: {}".format(epoch + 1, elbo.detach()))
    print("training takes {}s".format(time.time() - ts))

    if save_model:
        torch.save({'epoch': epoch, 'model_state_dict': model.state_dict(),
                    'optimizer_state_dict': optimizer.state_dict(), 'loss': loss}, PATH)
    if show_ELBO:
        elbo = model.compute_ELBO(X_list, Y_list)
        print("epoch: {}, ELBO: {}".format(epoch + 1, elbo.detach()))
    if X_test_list is not None:
        return model, loss_list, rmse_test_list, time_list
    return model, loss_list, time_list


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
