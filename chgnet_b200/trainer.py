"""Training step on the kernel engine (reference chgnet/trainer/trainer.py).

Mirrors the pieces of the reference ``Trainer`` that sit on the hot path of a fine-tuning step
(trainer.py:398-411): prediction -> ``CombinedLoss`` (779-869) -> ``loss.backward()`` ->
``optimizer.step()``, with the same constructor vocabulary (``targets``, ``criterion``,
``energy_loss_ratio`` ..., ``optimizer="Adam"``, ``learning_rate``, ``delta``,
``allow_missing_labels``).

What is covered: losses on energies and magnetic moments (``targets`` "e" / "em"),
MSE / MAE / Huber with NaN-masked missing labels, Adam as one fused kernel over a flat parameter
buffer, data-parallel training with ONE all-reduce of the flat gradient buffer per step
(SURVEY.md §8e).  Losses on forces / stresses need the second-order reverse pass, which is not
built yet: ``targets`` containing "f" or "s" raise ``NotImplementedError`` instead of training on
a silently incomplete gradient.

The loss normalisation follows the reference exactly for one process (``nn.MSELoss`` means over
the batch); across ranks the means are over the GLOBAL batch: the per-term numerators and counts
are all-reduced before the seeds are formed, and the gradient buffers are summed.
"""
from __future__ import annotations

from collections.abc import Sequence
from dataclasses import dataclass

import torch
from torch import Tensor

from chgnet_b200.weights import unpack_grads

_KIND = {"MSE": 0, "mse": 0, "MAE": 1, "mae": 1, "l1": 1, "Huber": 2}


@dataclass
class LossConfig:
    target_str: str = "e"
    criterion: str = "MSE"
    energy_loss_ratio: float = 1.0
    force_loss_ratio: float = 1.0
    stress_loss_ratio: float = 0.1
    mag_loss_ratio: float = 0.1
    delta: float = 0.1

    def __post_init__(self) -> None:
        if self.criterion not in _KIND:
            raise NotImplementedError(self.criterion)  # same as trainer.py:763
        if "f" in self.target_str or "s" in self.target_str:
            raise NotImplementedError(
                "chgnet_b200 trains on energy / magmom losses (targets 'e' or 'em'); force and stress losses "
                "need the second-order reverse pass, which is not built yet (DESIGN.md §9)")


def _all_reduce(t: Tensor, group) -> None:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)


def loss_and_seeds(K, cfg: LossConfig, e_pred: Tensor, e_target: Tensor, m_pred: Tensor | None,
                   m_target: Tensor | None, group=None) -> tuple[dict, Tensor, Tensor | None]:
    """CombinedLoss (trainer.py:779-869) on the device.

    Returns (report, dL/de [B], dL/dm [N] | None); ``report`` holds python floats: loss, e_MAE,
    e_MAE_size, (m_MAE, m_MAE_size).  NaN targets are missing labels.  With an initialised process
    group the means are over the global batch.
    """
    kind = _KIND[cfg.criterion]
    dev = e_pred.device
    sums = torch.zeros(2, 3, dtype=torch.float64, device=dev)
    g_e = torch.empty_like(e_pred)
    K.loss_terms(e_pred.contiguous(), e_target.contiguous(), kind, cfg.delta, g_e, sums[0])
    use_m = "m" in cfg.target_str and m_pred is not None and m_target is not None
    g_m = None
    if use_m:
        g_m = torch.empty_like(m_pred)
        K.loss_terms(m_pred.contiguous(), m_target.contiguous(), kind, cfg.delta, g_m, sums[1])
    _all_reduce(sums, group)
    cnt = sums[:, 2].clamp_min(1.0)
    g_e = g_e * (cfg.energy_loss_ratio / cnt[0]).to(g_e.dtype)
    if use_m:
        g_m = g_m * (cfg.mag_loss_ratio / cnt[1]).to(g_m.dtype)
    host = sums.cpu()
    n_e, n_m = max(float(host[0, 2]), 1.0), max(float(host[1, 2]), 1.0)
    report = {"loss": cfg.energy_loss_ratio * float(host[0, 0]) / n_e, "e_MAE": float(host[0, 1]) / n_e,
              "e_MAE_size": int(host[0, 2])}
    if "m" in cfg.target_str:
        report["m_MAE"], report["m_MAE_size"] = float(host[1, 1]) / n_m, int(host[1, 2])
        report["loss"] += cfg.mag_loss_ratio * float(host[1, 0]) / n_m
    return report, g_e, g_m


def loss_and_grads(engine, batch, cfg: LossConfig, e_target: Tensor, m_target: Tensor | None,
                   is_intensive: bool = True, group=None) -> tuple[dict, dict]:
    """One forward + loss + training reverse pass on an already built batch.

    Returns (report, packed-layout gradients of THIS rank's graphs under the global-batch loss);
    sum the gradients over ranks to get the global gradient.
    """
    K = engine.K
    out = engine.run(batch, need_grad=True, need_magmom="m" in cfg.target_str, train=True)
    n = torch.tensor(batch.atoms_per_graph, device=out.energy.device, dtype=out.energy.dtype)
    total = out.energy + out.e_ref
    e_pred = (total / n if is_intensive else total).to(out.site_e.dtype)
    report, g_e, g_m = loss_and_seeds(K, cfg, e_pred, e_target.to(e_pred.dtype), out.magmom,
                                      None if m_target is None else m_target.to(e_pred.dtype), group)
    seed_e = g_e / n.to(g_e.dtype) if is_intensive else g_e
    return report, engine.param_grads(out, seed_e.contiguous(), g_m)


class Trainer:
    """Fine-tuning loop on the CUDA kernel path (reference Trainer, trainer.py:37-411).

    >>> trainer = Trainer(model, targets="em", optimizer="Adam", criterion="MSE", learning_rate=1e-3)
    >>> report = trainer.train_step(graphs, {"e": e_labels, "m": [m_0, m_1, ...]})
    """

    def __init__(self, model, *, targets: str = "e", energy_loss_ratio: float = 1, force_loss_ratio: float = 1,
                 stress_loss_ratio: float = 0.1, mag_loss_ratio: float = 0.1, optimizer: str = "Adam",
                 criterion: str = "MSE", learning_rate: float = 1e-3, weight_decay: float = 0.0,
                 betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, delta: float = 0.1,
                 process_group=None, **_: object) -> None:
        if optimizer != "Adam":
            raise NotImplementedError("chgnet_b200.Trainer implements optimizer='Adam' (the reference default)")
        self.model = model
        self.cfg = LossConfig(targets, criterion, energy_loss_ratio, force_loss_ratio, stress_loss_ratio,
                              mag_loss_ratio, delta)
        self.lr, self.weight_decay, self.betas, self.eps = learning_rate, weight_decay, betas, eps
        self.group = process_group
        self.step_count = 0
        # one flat fp32 buffer holds every trainable parameter; the nn.Parameters become views of it,
        # so state_dict() stays the single source of truth and Adam is ONE kernel launch
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        self.names = [n for n, _ in named]
        self.shapes = [tuple(p.shape) for _, p in named]
        self.sizes = [p.numel() for _, p in named]
        # every parameter starts on a 64-byte boundary (the kernels read weights as float4); the
        # padding stays zero under Adam (zero gradient, zero moments)
        self.offsets, off = [], 0
        for sz in self.sizes:
            self.offsets.append(off)
            off += (sz + 15) // 16 * 16
        dev = named[0][1].device
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for (_, p), o, sz in zip(named, self.offsets, self.sizes):
            self.flat[o:o + sz] = p.detach().reshape(-1)
            p.data = self.flat[o:o + sz].view(p.shape)
        self.flat_grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        model.mark_params_updated()

    # ------------------------------------------------------------------
    def _targets(self, targets: dict, n_list: Sequence[int], device) -> tuple[Tensor, Tensor | None]:
        e_t = torch.as_tensor(targets["e"], dtype=torch.float32).reshape(-1).to(device)
        m_t = None
        if "m" in self.cfg.target_str:
            parts = []
            for n, m in zip(n_list, targets["m"]):  # None / NaN = missing labels (trainer.py:846-851)
                parts.append(torch.full((n,), float("nan")) if m is None else torch.as_tensor(m, dtype=torch.float32).reshape(-1))
            m_t = torch.cat(parts).to(device)
        return e_t, m_t

    def flatten_grads(self, grads: dict[str, Tensor]) -> Tensor:
        for name, o, sz in zip(self.names, self.offsets, self.sizes):
            self.flat_grad[o:o + sz] = grads[name].reshape(-1)
        return self.flat_grad

    def grads_by_name(self) -> dict[str, Tensor]:
        """views of the (all-reduced) flat gradient buffer, one per parameter"""
        return {n: self.flat_grad[o:o + sz].view(sh) for n, o, sz, sh in zip(self.names, self.offsets, self.sizes, self.shapes)}

    def train_step(self, graphs, targets: dict) -> dict:
        """prediction -> CombinedLoss -> parameter gradients -> (all-reduce) -> Adam; returns the report."""
        from chgnet_b200.batch import build_batch

        model = self.model
        engine = model._get_engine()
        compact = not any(gp.extra["bo"] is not None for gp in engine.pw.bond)
        batch = build_batch(graphs, model.device, with_reverse=True, compact_bonds=compact)
        e_t, m_t = self._targets(targets, batch.atoms_per_graph, model.device)
        report, G = loss_and_grads(engine, batch, self.cfg, e_t, m_t, model.is_intensive, self.group)
        flat_grad = self.flatten_grads(unpack_grads(G, model.state_dict()))
        _all_reduce(flat_grad, self.group)  # the one collective of the step
        self.step_count += 1
        engine.K.adam_step(self.flat, flat_grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0],
                           self.betas[1], self.eps, self.weight_decay, self.step_count)
        model.mark_params_updated()
        return report

    def train(self, loader, epochs: int = 1) -> list[dict]:
        """``loader`` yields (graphs, targets) like the reference's collate_graphs batches
        (dataset.py:763-788)."""
        history = []
        for _ in range(epochs):
            for graphs, targets in loader:
                history.append(self.train_step(graphs, targets))
        return history
