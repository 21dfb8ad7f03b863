"""N > 1 host logic (window sharding + score gather) on CPU: world_size-2 gloo processes, 127.0.0.1 rendezvous."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from plantcaduceus_b200.sharding import gather_rows, shard_range, score_sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 185, 256, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_score(ascii_rows: np.ndarray) -> np.ndarray:
    """Deterministic stand-in for the engine: a function of the window bytes only (so any mis-ordering shows)."""
    a = ascii_rows.astype(np.float32)
    return np.stack([a.sum(1), a[:, 0], a[:, -1], (a * np.arange(a.shape[1])).sum(1)], axis=1).astype(np.float32)


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        windows = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=(n, 64))
        full = score_sharded(_fake_score, windows)
        ok = np.array_equal(full, _fake_score(windows))
        # gather_rows with uneven shards
        s, e = shard_range(n, rank, world)
        g = gather_rows(torch.arange(s, e, dtype=torch.float32)[:, None], n)
        ok = ok and torch.equal(g[:, 0], torch.arange(n, dtype=torch.float32))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [185, 2, 1])
def test_score_gather_world_size_2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(2))
    assert results == {0: True, 1: True}


class _FakeEngine:
    """CPU stand-in with the two device entry points genome_scan uses: windows by the host rule, scores = a function of
    the window bytes (so a wrong chromosome, position or order shows)."""
    device = torch.device("cpu")

    def extract_windows_device(self, chrom_dev, pos0, token_idx=255, length=512):
        from plantcaduceus_b200 import genome_io as gio
        chrom = bytes(chrom_dev.numpy())
        rows = [np.frombuffer(gio.extract_window(chrom, int(p), token_idx, length), dtype=np.uint8) for p in pos0.tolist()]
        return torch.from_numpy(np.stack(rows).copy()) if rows else torch.zeros((0, length), dtype=torch.uint8)

    def score_windows_device(self, windows, token_idx, out=None):
        res = torch.from_numpy(_fake_score(windows.numpy()))
        if out is not None:
            out.copy_(res)
            return out
        return res


def _genome_case():
    rng = np.random.default_rng(5)
    seqs = {f"chr{i}": bytes(rng.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8), size=700 + 300 * i)) for i in range(3)}
    names = ["chr2", "chr0", "chr1"]          # order of first appearance in the VCF, not the FASTA's
    cid = rng.integers(0, 2, 41).astype(np.int32)       # chr1 (id 2) carries no variant: never broadcast
    pos = np.array([rng.integers(0, len(seqs[names[c]])) for c in cid], dtype=np.int64)
    return names, seqs, cid, pos


def _worker_genome(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from plantcaduceus_b200 import genome_io as gio
        from plantcaduceus_b200.genome_scan import score_variants_sharded
        from plantcaduceus_b200.sharding import scatter_rows
        names, seqs, cid, pos = _genome_case()
        args = (names, seqs, cid, pos) if rank == 0 else (None, None, None, None)     # only rank 0 has parsed anything
        got = score_variants_sharded(_FakeEngine(), *args, batch_size=7, token_idx=20, length=64, device=torch.device("cpu"))
        want = _fake_score(np.stack([np.frombuffer(gio.extract_window(seqs[names[c]], int(p), 20, 64), dtype=np.uint8)
                                     for c, p in zip(cid, pos)]))
        ok = np.array_equal(got, want)
        rows = torch.arange(23 * 3, dtype=torch.uint8).reshape(23, 3) if rank == 0 else None
        mine = scatter_rows(rows)
        s, e = shard_range(23, rank, world)
        ok = ok and torch.equal(mine, torch.arange(23 * 3, dtype=torch.uint8).reshape(23, 3)[s:e])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_rank0_parses_and_broadcasts_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_genome, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(2))
    assert results == {0: True, 1: True}


def test_score_variants_single_process_and_fasta_reader(tmp_path):
    from plantcaduceus_b200 import genome_io as gio
    from plantcaduceus_b200.genome_scan import score_variants_sharded
    names, seqs, cid, pos = _genome_case()
    got = score_variants_sharded(_FakeEngine(), names, seqs, cid, pos, batch_size=16, token_idx=20, length=64, device=torch.device("cpu"))
    want = _fake_score(np.stack([np.frombuffer(gio.extract_window(seqs[names[c]], int(p), 20, 64), dtype=np.uint8)
                                 for c, p in zip(cid, pos)]))
    assert np.array_equal(got, want)
    # vectorised FASTA reader: wrapped lines, CRLF, blank lines, description after the id, text before the first header
    fa = tmp_path / "g.fa"
    body = lambda s, w, nl: nl.join(s[i:i + w].decode() for i in range(0, len(s), w))
    fa.write_bytes(("junk before any header\n>chr0 first record\n" + body(seqs["chr0"], 60, "\n") + "\n\n>chr1\tdesc\r\n"
                    + body(seqs["chr1"], 50, "\r\n") + "\r\n>chr2\n" + body(seqs["chr2"], 80, "\n")).encode())
    assert gio.read_fasta(str(fa)) == seqs
    import gzip
    with gzip.open(str(fa) + ".gz", "wb") as f:
        f.write(fa.read_bytes())
    assert gio.read_fasta(str(fa) + ".gz") == seqs
    fa.write_bytes(b">a\nAC\n>a\nGT\n")
    with pytest.raises(ValueError, match="duplicate"):
        gio.read_fasta(str(fa))


def _worker_cli(rank, world, ports, tmp, q):
    """The command line itself under two gloo ranks, GPU steps stubbed (fake engine): rank 0 alone reads the inputs,
    windows / coordinates travel over the process group, rank 0 writes; the other rank returns without output."""
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(ports[0]))
    from plantcaduceus_b200 import CharDNATokenizer
    from plantcaduceus_b200 import genome_io as gio
    from plantcaduceus_b200 import zero_shot_score as zss
    tok = CharDNATokenizer()

    class Engine(_FakeEngine):                       # bounded "logits", so that the scores are finite numbers
        def score_windows_device(self, windows, token_idx, out=None):
            res = 3 * torch.sin(super().score_windows_device(windows, token_idx))
            return res if out is None else out.copy_(res)

    zss.load_model_and_tokenizer = lambda *a, **k: (Engine(), tok)

    def fake_extract_logits(model, dataloader, device, tokenIdx, tokenizer):
        dataset, batch_size = dataloader
        out = np.zeros((len(dataset), 4), dtype=np.float32)
        for start, batch in dataset.ascii_batches(batch_size):
            out[start:start + len(batch)] = 3 * np.sin(_fake_score(np.asarray(batch)))
        return gio.softmax4(out)

    zss.extract_logits = fake_extract_logits
    gold = os.path.join(ROOT, "tests", "golden")
    rc1 = zss.main(["-input-table", os.path.join(gold, "example_snp.tsv"), "-output", os.path.join(tmp, f"w{world}.tsv"),
                    "-device", "cpu", "-batchSize", "32"])
    os.environ["MASTER_PORT"] = str(ports[1])          # main() tears its process group down: a fresh rendezvous per call
    rc2 = zss.main(["-input-vcf", os.path.join(gold, "example_maize_snp.vcf"), "-input-fasta",
                    os.path.join(gold, "example_genome.fa.gz"), "-output", os.path.join(tmp, f"w{world}.vcf"),
                    "-device", "cpu", "-batchSize", "32"])
    q.put((rank, rc1 == 0 and rc2 == 0))


def test_cli_world_size_2_equals_single_process(tmp_path):
    """python -m plantcaduceus_b200.zero_shot_score under 2 ranks writes byte-identical files to the 1-process run
    (table path: packed windows scattered; VCF path: coordinates + chromosomes broadcast; scores gathered once)."""
    ctx = mp.get_context("spawn")
    for world in (1, 2):
        q = ctx.Queue()
        ports = (_free_port(), _free_port())
        procs = [ctx.Process(target=_worker_cli, args=(r, world, ports, str(tmp_path), q)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(240)
            if p.is_alive():
                p.kill()
            assert p.exitcode == 0
        assert dict(q.get(timeout=10) for _ in range(world)) == {r: True for r in range(world)}
    for ext in ("tsv", "vcf"):
        one, two = (tmp_path / f"w1.{ext}").read_bytes(), (tmp_path / f"w2.{ext}").read_bytes()
        assert len(one) > 1000 and one == two
    text = (tmp_path / "w1.vcf").read_text()
    assert text.count("plantCAD_zero_shot=") == 190 and "inf" not in text and "nan" not in text
