"""Host-side mirror of `salmon quant` for the hot path (src/quant/SalmonQuantify.cpp:2777-2956 ->
quantifyLibrary -> stageFinalizeMappingOutputs, src/quant/pipeline/MappingPipelineStages.cpp:37-206):

    reads --sb_map_batch--> per-read alignments, online masses / FLD, per-batch class tables
          --sb_map_finish--> equivalence classes (EquivalenceClassBuilder::finish) + normalizeAlphas
          --sb_em_optimize--> alpha (CollapsedEMOptimizer::optimize)
          --sb_tpm / sb_write_quant_sf / sb_write_eq_classes--> quant.sf, aux_info/eq_classes.txt.gz

One process per GPU; with torch.distributed initialised (world > 1) every rank maps its own read shard, the
end-of-mapping statistics are reduced once (salmon_b200.dist) and the EM all-reduces alpha per iteration (NCCL).
"""
from __future__ import annotations

import os

import numpy as np

from . import _capi
from ._capi import EMContext, EqClasses, MapContext, default_params, map_default_params


def quant_reads(index, left, right, map_params=None, em_params=None, device=0, batch=262_144, dist=None,
                names=None, out_dir=None, dump_eq=False, dump_eq_weights=False):
    """left/right: [n, L] uint8 base codes (0..3 = ACGT, 4 = N) of THIS rank's read shard.
    Returns dict(alpha, tpm, eff_len, classes, n_mapped, em_stats)."""
    mp = map_params or map_default_params()
    ep = em_params or default_params()
    n, L = left.shape
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    ctx = MapContext(index, mp, device=device, batch_cap=min(batch, max(n, 1)), max_read_len=L)
    for s in range(0, n, batch):
        ctx.map_batch(left[s:s + batch], right[s:s + batch])
    res = ctx.finish()
    inputs = res
    n_mapped = res["counters"]["n_mapped"]
    if world > 1:
        from .dist import reduce_partials
        g, roots = reduce_partials(ctx.partial(), dist, f"cuda:{device}")
        inputs = ctx.project_global(g, roots)
        n_mapped = g["assigned"]
    ctx.close()
    M = index.n_txps
    eq = EqClasses(M, res["off"], res["tids"], res["weights"], res["counts"])
    em = EMContext(device)
    if world > 1:
        em.peer_setup(dist, M)          # alpha all-reduced inside the persistent kernel over NVLink peer memory
    alpha, st, ok = em.optimize(eq, ep, inputs["projected_counts"], inputs["eff_len"], inputs["unique_counts"])
    em.close()
    if not ok:
        raise _capi.SalmonB200Error("The optimization algorithm failed (total alpha weight too small)")
    tpm = _capi.tpm(alpha, inputs["eff_len"], float(n_mapped) if n_mapped else None)
    if out_dir is not None and rank == 0:
        os.makedirs(os.path.join(out_dir, "aux_info"), exist_ok=True)
        nm = names or [f"t{i}" for i in range(M)]
        lens = (index.off[1:] - index.off[:-1]).astype(np.uint32)
        _capi.write_quant_sf(os.path.join(out_dir, "quant.sf"), nm, lens, inputs["eff_len"], alpha,
                             float(n_mapped) if n_mapped else None)
        if dump_eq or dump_eq_weights:
            _capi.write_eq_classes(os.path.join(out_dir, "aux_info", "eq_classes.txt.gz"), nm, res["off"], res["tids"],
                                   res["counts"], res["weights"] if dump_eq_weights else None)
    return dict(alpha=alpha, tpm=tpm, eff_len=inputs["eff_len"], classes=res, n_mapped=int(n_mapped), em_stats=st,
                projected_counts=inputs["projected_counts"], unique_counts=inputs["unique_counts"])
