"""Run the BASELINE.json configurations at (or near) full size and print one JSON
record per configuration.  Single process, or one process per GPU under torchrun:

    python tools/run_configs.py --configs 1,2,3,4 [--scale 1.0]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        tools/run_configs.py --configs 3,5

`--scale s` multiplies the *query / output* counts (not the list sizes) so a
configuration can be bounded to the GPU-minutes available; every record states
the sizes it actually ran.  Timings are wall-clock around the public API call
(host arrays in, host arrays out), max over ranks.  Parity spot checks use the
CPU oracle on a bounded sample (test infrastructure, not the measured path).
"""
import argparse
import contextlib
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

from bench import synthetic_scaled  # noqa: E402
from kdsource_b200 import _lib, dist, kde  # noqa: E402
from kdsource_b200.sampler import DeviceSampler, make_geom_struct  # noqa: E402
from kdsource_b200.treekde import TreeKDE  # noqa: E402


def sync_max(dt):
    torch.cuda.synchronize()
    rank, world = dist.rank_world()
    if world == 1:
        return dt
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def sum_ranks(x):
    rank, world = dist.rank_world()
    if world == 1:
        return float(x)
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return float(t.item())


def timed(fn):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, sync_max(time.perf_counter() - t0)


def mem_available_gb():
    try:
        with open("/proc/meminfo") as fh:
            for ln in fh:
                if ln.startswith("MemAvailable"):
                    return int(ln.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def config1(args):
    """Verification example: 1e5-particle list, Silverman bw, evaluate on a 1e4 grid,
    through KDSource (MCPL file -> fit -> evaluate)."""
    import kdsource_b200.api as kds
    from kdsource_b200 import mcpl
    from oracle import kds_oracle as O

    rank, world = dist.rank_world()
    parts, w, g, _, _ = synthetic_scaled(100_000)
    qparts = synthetic_scaled(10_000, seed=54321, wseed=1)[0]
    tmp = tempfile.mkdtemp(prefix="kdsb_c1_")
    try:
        fn = os.path.join(tmp, "list{}.mcpl.gz".format(rank))
        mcpl.write_mcpl_file(fn, ekin=parts[:, 0], x=parts[:, 1], y=parts[:, 2], z=parts[:, 3],
                             ux=parts[:, 4], uy=parts[:, 5], uz=parts[:, 6], time=parts[:, 7],
                             weight=w, pdgcode=2112, srcname="synthetic")
        s = kds.KDSource(kds.PList(fn), g, bw="silv")
        _, t_fit = timed(lambda: s.fit())
        s.evaluate(qparts)  # warm-up (context, source packing, buffers of this shape)
        (ev, er), t_eval = timed(lambda: s.evaluate(qparts))
        fparts, fw = s.plist.get()
        # the notebook's bandwidth optimisation (Verification.ipynb cells 13-15): adaptive
        # kNN seed + MLCV rescale on the first 1e4 particles (stored output: 14.6-15.0 s)
        s2 = kds.KDSource(kds.PList(fn), g, bw="auto")
        _, t_auto = timed(lambda: s2.fit(Nmlcv=1e4, Nbatch=1e4, Nknn=10, show=False,
                                         grid=np.logspace(-0.8, 0.8, 33)))
        bw_auto_stats = [float(np.min(s2.kde.bw)), float(np.median(s2.kde.bw)),
                         float(np.max(s2.kde.bw))]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    vecs = g.transform(fparts.copy()) / s.scaling
    qv = g.transform(qparts.copy()) / s.scaling
    ref = O.kde_sum_c(vecs, fw, s.kde.bw, qv) / np.prod(s.scaling) * g.jac(qparts.copy())
    ok = ref > 1e-12 * ref.max()
    return {"config": 1, "N": 100_000, "M": 10_000, "bw": float(s.kde.bw), "fit_s": t_fit,
            "evaluate_s": t_eval, "pairs_per_s_e2e": 1e9 / t_eval,
            "fit_bw_auto_s": t_auto, "bw_auto_min_med_max": bw_auto_stats,
            "max_rel_err_vs_oracle": float(np.max(np.abs(ev[ok] / ref[ok] - 1)))}


def config2(args):
    """MLCV bandwidth optimisation, 1e6-particle 6-D list, 32-point grid."""
    N = 1_000_000
    _, w, _, _, data = synthetic_scaled(N)
    seed = kde.bw_silv(6, np.sum(w) ** 2 / np.sum(w ** 2))
    # the optimum of this list sits below Silverman's rule: centre the grid there
    grid = np.logspace(-0.75, -0.15, 32)
    kde.mlcv_scores(data[:20000], w[:20000], seed, grid)  # warm-up
    scores, t = timed(lambda: kde.mlcv_scores(data, w, seed, grid, n_splits=10))
    fb = kde.kfold_bounds(N, 10)
    nf = np.diff(fb).astype(np.float64)
    evals = float(np.sum(nf * (N - nf))) * len(grid)
    best = int(np.argmax(scores))
    raised = None
    try:
        bw, _ = timed(lambda: kde.bw_mlcv(data, w, 10, seed, grid, show=False))
        bw = float(bw)
    except Exception as e:  # reference behaviour when the maximum is at an edge
        raised, bw = str(e), None
    return {"config": 2, "N": N, "G": len(grid), "K": 10, "mlcv_scores_s": t,
            "evals_per_s_e2e": evals / t, "argmax": best, "bw_mlcv": bw,
            "bw_silv": float(seed), "raised": raised,
            "scores_first_best_last": [float(scores[0]), float(scores[best]), float(scores[-1])]}


def config3(args):
    """kNN adaptive bandwidth (k=10, batches of 1e4) on a 1e7-particle list, then
    evaluate at 1e8 query points (each rank draws and evaluates its own share)."""
    from oracle import kds_oracle as O

    rank, world = dist.rank_world()
    N = 10_000_000
    M_tot = int(1e8 * args.scale)
    _, w, _, _, data = synthetic_scaled(N)
    bw, t_knn = timed(lambda: kde.bw_knn(data, w, k=10, batch_size=10000))
    off = kde.array_split_bounds(N, 1000)
    knn_pairs = float(np.sum(np.diff(off).astype(np.float64) ** 2))
    lo, hi = dist.shard_range(M_tot, rank, world)
    q = synthetic_scaled(hi - lo, seed=54321 + rank, wseed=1)[4]
    k = TreeKDE(bw=bw).fit(data, w)
    k.evaluate(q[:1024], shard=False)  # pack + upload the sources, first launch
    dens, t_eval = timed(lambda: k.evaluate(q, shard=False))
    checksum = sum_ranks(float(np.sum(dens)))
    rec = {"config": 3, "N": N, "k": 10, "batch_size": 10000, "knn_s": t_knn,
           "knn_pairs_per_s_e2e": knn_pairs / t_knn, "M": M_tot, "evaluate_s": t_eval,
           "pairs_per_s_e2e": float(N) * M_tot / t_eval, "density_checksum": checksum,
           "bw_min_med_max": [float(bw.min()), float(np.median(bw)), float(bw.max())]}
    if rank == 0:
        ns = 2000
        ref = O.kde_sum_c(data, w, bw, q[:ns], nthreads=os.cpu_count() or 1)
        ok = ref > 1e-12 * ref.max()
        rec["max_rel_err_vs_oracle_{}q".format(ns)] = float(
            np.max(np.abs(dens[:ns][ok] / ref[ok] - 1)))
        sub = slice(0, 20000)
        rec["knn_bit_exact_first_2_batches"] = bool(
            np.array_equal(bw[sub], O.bw_knn(data[sub], w[sub], k=10, batch_size=10000)))
    return rec


def config4(args):
    """kdsource-resample: new particles from a 1e7-particle adaptive-bw source to MCPL.
    (a) device rate for 1e9 particles (records produced in HBM, 1e8 per launch);
    (b) resample_to_mcpl end to end to a gzip MCPL file for 1e8*scale particles."""
    import kdsource_b200.api as kds
    from kdsource_b200 import mcpl

    rank, world = dist.rank_world()
    N = 10_000_000
    ps, w, g, sc, _ = synthetic_scaled(N, seed=7)
    bws = np.random.default_rng(3).uniform(0.2, 0.5, N).astype(np.float32)
    smp, t_setup = timed(lambda: DeviceSampler(ps, w, bws, make_geom_struct(g, sc, "g")))
    n_dev = 1_000_000_000
    lo, hi = dist.shard_range(n_dev, rank, world)
    per = 100_000_000
    out = _lib.empty(per * 36, torch.uint8)

    def gen():
        for s in range(lo, hi, per):
            smp.sample_device(min(per, hi - s), seed=11, first=s, out=out)

    gen()
    _, t_dev = timed(gen)
    del out
    rec = {"config": 4, "N": N, "sampler_setup_s": t_setup, "device_particles": n_dev,
           "device_s": t_dev, "device_particles_per_s": n_dev / t_dev}
    nout = int(1e8 * args.scale)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = os.path.join(base or tempfile.gettempdir(), "kdsb_c4")
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp)
        mcpl.write_mcpl_file(os.path.join(tmp, "list.mcpl.gz"), ekin=ps[:, 0], x=ps[:, 1],
                             y=ps[:, 2], z=ps[:, 3], ux=ps[:, 4], uy=ps[:, 5], uz=ps[:, 6],
                             time=ps[:, 7], weight=w, pdgcode=2112, srcname="synthetic")
    dist.barrier()
    cwd = os.getcwd()
    try:
        os.chdir(tmp)
        src = kds.KDSource(kds.PList("list.mcpl.gz", set_params=False), g, bw=bws.astype(np.float64))
        if rank == 0:
            src.scaling = np.asarray(sc, dtype=np.float64)  # bandwidths given: no fit needed to save
            src.save("src.xml", adjust_N=False)
        dist.barrier()
        from kdsource_b200.resample import open_source

        _, t_open = timed(lambda: open_source("src.xml"))
        rec["open_source_s"] = t_open
        _, t_e2e = timed(lambda: kds.resample_to_mcpl("src.xml", "out.mcpl.gz", nout, rng=1,
                                                      force=True))
        if rank == 0:
            rec.update(e2e_particles=nout, e2e_s=t_e2e, e2e_particles_per_s=nout / t_e2e,
                       e2e_file_bytes=os.path.getsize("out.mcpl.gz"),
                       e2e_what="resample_to_mcpl: XML + 1e7-particle MCPL read and device decode, "
                                "generation, device gzip, D2H, file on " + (base or "tmp"))
            with mcpl.MCPLFile("out.mcpl.gz", blocklength=100000) as f:
                pb = f.read_block()
                rec["e2e_file_nparticles"] = int(f.nparticles)
                rec["e2e_first_block_ok"] = bool(np.all(pb.weight == 1) and np.all(pb.z == 0)
                                                 and np.all(pb.pdgcode == 2112))
    finally:
        os.chdir(cwd)
        dist.barrier()
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)
    return rec


def config5(args):
    """Full pipeline on a 1e8-particle list, as a user would run it: an MCPL file on disk ->
    KDSource.fit with the kNN + MLCV bandwidth (bw="auto"), device-resident and sharded over
    the ranks -> a density map of 1e9*scale query points (hard support radius proportional to
    each source's bandwidth, tiles beyond it culled exactly) -> 1e10 resamples.  Query points
    are jittered copies of list entries drawn on the device (they follow the data, where the
    work is); every rank evaluates its share in blocks of 2^25."""
    import kdsource_b200.api as kds
    from kdsource_b200 import mcpl

    rank, world = dist.rank_world()
    N = int(args.n5)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    tmp = os.path.join(base, "kdsb_c5")
    fn = os.path.join(tmp, "list.mcpl")
    # the list: every rank draws its share and writes its records into place
    t0 = time.perf_counter()
    lo, hi = dist.shard_range(N, rank, world)
    ps, w, g, _, _ = synthetic_scaled(hi - lo, seed=5 + 1000 * rank, wseed=777 + rank)
    rec, hdr = mcpl.pack_records(ekin=ps[:, 0], x=ps[:, 1], y=ps[:, 2], z=ps[:, 3], ux=ps[:, 4],
                                 uy=ps[:, 5], uz=ps[:, 6], time=ps[:, 7], weight=w, pdgcode=2112,
                                 srcname="synthetic", nparticles=N)
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp)
        with open(fn, "wb") as fh:
            fh.write(hdr)
            fh.truncate(len(hdr) + N * rec.dtype.itemsize)
    dist.barrier()
    with open(fn, "r+b") as fh:
        fh.seek(len(hdr) + lo * rec.dtype.itemsize)
        fh.write(rec.tobytes())
    del rec, ps, w
    dist.barrier()
    t_list = time.perf_counter() - t0
    note = None
    try:
        s = kds.KDSource(kds.PList(fn, set_params=False), g, bw="auto")
        # wide, and centred below 1: the kNN seed (10th neighbour inside batches of 1e4) is
        # several times the likelihood-optimal bandwidth of a list this dense
        grid = np.logspace(-1.4, 0.2, 32)

        def fit():
            try:
                s.fit(device=True, sharded=world > 1, Nmlcv=1e6, Nbatch=1e4, Nknn=10, show=False,
                      grid=grid)
            except Exception as e:  # maximum at a grid edge: keep the kNN seed, report it
                if "Maximum not found" not in str(e):
                    raise
                print("bw_auto:", e)
                s.fit(device=True, sharded=world > 1, bw_method="knn", k=10, batch_size=10000)
                return str(e)
            return None

        note, t_fit = timed(fit)
        k = s.kde
        k.cutoff = "adaptive"
        _, t_pack = timed(lambda: k._device_state())
        dev = k._device_state()
        # density map: jittered copies of list entries, generated on the device
        M_tot = int(1e9 * args.scale)
        qlo, qhi = dist.shard_range(M_tot, rank, world)
        blk = 1 << 25
        gen = torch.Generator(device="cuda")
        gen.manual_seed(4242 + rank)
        d_data = k._d_data
        bw_typ = float(np.median(k.bw[::1000]))

        def density_map():
            visited, checksum, nz = 0, 0.0, 0
            for q0 in range(qlo, qhi, blk):
                m = min(blk, qhi - q0)
                idx = torch.randint(0, N, (m,), device="cuda", generator=gen)
                d_q = d_data[idx] + bw_typ * torch.randn((m, d_data.shape[1]), device="cuda",
                                                         dtype=torch.float64, generator=gen)
                out = k.evaluate_device(d_q, m, dev)
                visited += int(k.last_tiles_visited.item())
                checksum += float(out.sum().item())
                nz += int((out > 0).sum().item())
            return visited, checksum, nz

        (visited, checksum, nz), t_map = timed(density_map)
        executed = sum_ranks(visited * 512.0 * 1024.0)
        dense = float(N) * M_tot
        # spot check against the oracle's per-source hard radius on a few queries (rank 0)
        spot = None
        if rank == 0:
            from oracle import kds_oracle as O

            m = 32
            idx = torch.randint(0, N, (m,), device="cuda", generator=gen)
            d_q = d_data[idx] + bw_typ * torch.randn((m, 6), device="cuda", dtype=torch.float64,
                                                     generator=gen)
            got = k.evaluate_device(d_q, m, dev).cpu().numpy()
            nsub = min(N, 5_000_000)  # the oracle on a bounded prefix of the sources
            ksub = TreeKDE(bw=k.bw[:nsub], cutoff="adaptive").fit(k.data[:nsub], k.weights[:nsub])
            gsub = ksub.evaluate(d_q.cpu().numpy(), shard=False)
            # the radius factor follows the smallest bandwidth of the fitted set: take the
            # subset model's own
            c = ksub._device_state()["cull_rel"]
            ref = O.kde_sum(k.data[:nsub], k.weights[:nsub], k.bw[:nsub], d_q.cpu().numpy(),
                            cutoff_rel=c, chunk=4)
            ok = ref > 1e-9 * ref.max()
            spot = {"queries": m, "sources": nsub,
                    "max_rel_err_vs_oracle": float(np.max(np.abs(gsub[ok] / ref[ok] - 1)))}
            del ksub
        # resampling: records stay in HBM
        parts_d, ws_d = s.plist.get_device(-1)
        smp, t_setup = timed(lambda: DeviceSampler(parts_d, ws_d, k.bw.astype(np.float32),
                                                   make_geom_struct(g, s.scaling, "g")))
        del parts_d, ws_d
        n_res = 10_000_000_000
        rlo, rhi = dist.shard_range(n_res, rank, world)
        per = 125_000_000
        out = _lib.empty(per * 36, torch.uint8)

        def gen_res():
            for s0 in range(rlo, rhi, per):
                smp.sample_device(min(per, rhi - s0), seed=13, first=s0, out=out)

        _, t_res = timed(gen_res)
    finally:
        dist.barrier()
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)
    return {"config": 5, "N": N, "list_generation_and_write_s": t_list,
            "list_file": "uncompressed MCPL-3 on " + base, "fit_bw_auto_s": t_fit, "note": note,
            "N_eff": float(s.N_eff), "bw_min_med_max": [float(np.min(k.bw)), bw_typ,
                                                        float(np.max(k.bw))],
            "source_order_pack_s": t_pack, "tile_extent": dev["tile_extent"],
            "cutoff": "adaptive: r_i = {:.3f} h_i".format(dev["cull_rel"]),
            "density_queries": M_tot, "density_map_s": t_map,
            "dense_equivalent_pairs_per_s": dense / t_map,
            "executed_pairs_per_s": executed / t_map, "visited_fraction": executed / dense,
            "density_checksum": sum_ranks(checksum), "nonzero_queries": sum_ranks(nz),
            "spot_check": spot, "sampler_setup_s": t_setup, "resamples": n_res,
            "resample_s": t_res, "resampled_particles_per_s": n_res / t_res}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--n5", type=float, default=1e8, help="list size of configuration 5")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.rank_world()[0]
    fns = {1: config1, 2: config2, 3: config3, 4: config4, 5: config5}
    for c in [int(x) for x in args.configs.split(",")]:
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sys.stderr):
            rec = fns[c](args)
        rec.update(n_gpus=world, scale=args.scale, wall_s=time.perf_counter() - t0,
                   host_cores=os.cpu_count())
        if rank == 0:
            print(json.dumps(rec), flush=True)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
