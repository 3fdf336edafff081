"""PSTJ trajectory snapshots (include/ps_snapshot.h, physicssimulator_b200/snapshot.py — SURVEY.md §8(f) rank 4):
the format round-trips, its C header and the Python writer agree on every offset, the committed golden files are
reproduced by the oracle (CPU) and replayed bit for bit by the CUDA path through the C ABI (GPU)."""
import ctypes
import io
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from common import ROOT, same_bits
from golden.make_snapshots import CASES, generate
from physicssimulator_b200 import snapshot
from physicssimulator_b200.configuration import ps_step_config

GOLD = os.path.join(ROOT, "tests", "golden")


def _assert_same(a: snapshot.Trajectory, b: snapshot.Trajectory, what=""):
    assert bytes(a.config) == bytes(b.config)
    assert np.array_equal(a.world_offset, b.world_offset) and a.flags == b.flags
    for k, v in a.constants.items():
        assert same_bits(v, b.constants[k]) if v.dtype == np.float64 else np.array_equal(v, b.constants[k]), f"{what}: {k}"
    assert [f.step for f in a.frames] == [f.step for f in b.frames]
    for d in snapshot.compare(a, b, static_ang_mom=True):
        assert d.bit_exact and d.candidates_equal and d.pairs_equal and d.features_equal and d.contact_max_abs == 0.0, f"{what}: {d}"


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden_snapshot(name):
    gold = snapshot.load(os.path.join(GOLD, name + ".pstj"))
    assert gold.producer == snapshot.PRODUCER_ORACLE and gold.scene == name
    assert gold.flags == snapshot.HAS_CANDIDATES | snapshot.HAS_CONTACTS
    assert gold.frames[0].step == 0 and all(len(c["pairs"]) == 0 for c in gold.frames[0].contacts)
    assert any(len(c["pairs"]) > 0 for f in gold.frames[1:] for c in f.contacts), "a golden trajectory without one contact pins nothing"
    _assert_same(gold, generate(name), name)


def test_round_trip_is_byte_identical_and_truncation_is_loud():
    name = sorted(CASES)[-1]
    raw = open(os.path.join(GOLD, name + ".pstj"), "rb").read()
    t = snapshot.read(io.BytesIO(raw))
    assert snapshot.dumps(t) == raw
    with pytest.raises(snapshot.SnapshotError):
        snapshot.read(io.BytesIO(raw[:len(raw) - 9]))
    with pytest.raises(snapshot.SnapshotError):
        snapshot.read(io.BytesIO(b"XXXX" + raw[4:]))
    bad = bytearray(raw)
    struct.pack_into("<I", bad, 4, 99)  # version
    with pytest.raises(snapshot.SnapshotError):
        snapshot.read(io.BytesIO(bytes(bad)))


def test_compare_reports_a_perturbation_and_a_changed_set():
    name = sorted(CASES)[0]
    a = snapshot.load(os.path.join(GOLD, name + ".pstj"))
    b = snapshot.load(os.path.join(GOLD, name + ".pstj"))
    dyn = np.flatnonzero(~np.isinf(b.constants["mass"]))
    b.frames[2].state["pos"][dyn[0], 0] *= (1.0 + 1e-13)
    hit = next(w for w, c in enumerate(b.frames[3].contacts) if len(c["pairs"]))
    b.frames[3].contacts[hit]["feature"][0] ^= 1
    d = snapshot.compare(a, b)
    assert d[0].bit_exact and d[1].bit_exact
    assert not d[2].bit_exact and 0 < d[2].max_rel["pos"] < 1e-12 and d[2].max_rel["vel"] == 0.0
    assert d[3].pairs_equal and not d[3].features_equal


def test_c_header_and_python_writer_agree_on_the_layout():
    """compiles a 30-line C reader against include/ps_snapshot.h and lets it walk a golden file"""
    name = sorted(CASES)[-1]
    path = os.path.join(GOLD, name + ".pstj")
    t = snapshot.load(path)
    src = r'''
#include <stdio.h>
#include <stdlib.h>
#include "ps_snapshot.h"
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    ps_snapshot_header h;
    if (!f || fread(&h, sizeof h, 1, f) != 1) return 2;
    if (h.magic != PS_SNAP_MAGIC || h.version != PS_SNAP_VERSION || h.header_bytes != sizeof h || h.config_bytes != sizeof(ps_step_config)) return 3;
    int32_t* off = malloc(4 * (h.n_worlds + 1));
    if (fread(off, 4, h.n_worlds + 1, f) != h.n_worlds + 1) return 4;
    long n = h.n_bodies;
    fseek(f, n * (4 + 8 * (3 + 1 + 3 + 3 + 3 + 1 + 1 + 1)), SEEK_CUR);   /* constants */
    long pairs = 0, contacts = 0, cands = 0;
    for (uint32_t k = 0; k < h.n_frames; k++) {
        ps_snapshot_frame_header fh;
        if (fread(&fh, sizeof fh, 1, f) != 1 || fh.magic != PS_SNAP_FRAME_MAGIC) return 5;
        fseek(f, n * 8 * 18, SEEK_CUR);
        if (h.flags & PS_SNAP_HAS_CANDIDATES)
            for (uint32_t w = 0; w < h.n_worlds; w++) { int32_t c; if (fread(&c, 4, 1, f) != 1) return 6; cands += c; fseek(f, 8L * c, SEEK_CUR); }
        if (h.flags & PS_SNAP_HAS_CONTACTS)
            for (uint32_t w = 0; w < h.n_worlds; w++) {
                int32_t pc[2];
                if (fread(pc, 4, 2, f) != 2) return 7;
                pairs += pc[0]; contacts += pc[1];
                fseek(f, 8L * pc[0] + 4L * (pc[0] + 1) + 4L * pc[1] + 8L * 7 * pc[1], SEEK_CUR);
            }
    }
    long here = ftell(f); fseek(f, 0, SEEK_END);
    if (here != ftell(f)) return 8;
    printf("%u %u %u %d %.17g %ld %ld %ld %s\n", h.n_worlds, h.n_bodies, h.n_frames, h.config.solver_iterations, h.config.dt, cands, pairs, contacts, h.scene);
    return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "r.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "r")
        subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe, path], text=True).split()
    cands = sum(len(c) for f in t.frames for c in f.candidates)
    pairs = sum(len(c["pairs"]) for f in t.frames for c in f.contacts)
    conts = sum(len(c["feature"]) for f in t.frames for c in f.contacts)
    assert out == [str(t.n_worlds), str(t.n_bodies), str(len(t.frames)), str(t.config.solver_iterations), repr(t.config.dt), str(cands), str(pairs), str(conts), name]
    assert ctypes.sizeof(ps_step_config) == 128


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_replays_golden_snapshot(name):
    """the file alone (constants + frame 0 + config) rebuilds the batch; every later frame must come out bit for bit"""
    from physicssimulator_b200 import _native
    gold = snapshot.load(os.path.join(GOLD, name + ".pstj"))
    cfg = ps_step_config.from_buffer_copy(bytes(gold.config))
    cfg.record_contacts = 1
    batch = _native.Batch(gold.bodies(0), gold.world_offset, cfg, device=0)
    try:
        marks = [f.step for f in gold.frames[1:]]
        got = snapshot.record(gold.bodies(0), gold.world_offset, gold.config, marks, batch.step, batch.get_state,
                              batch.contacts, batch.candidates, producer=snapshot.PRODUCER_CUDA, scene=name)
    finally:
        batch.close()
    for d in snapshot.compare(gold, got):  # dynamic bodies; a static body's L is quirk Q2
        assert d.bit_exact and d.candidates_equal and d.pairs_equal and d.features_equal and d.contact_max_abs == 0.0, f"{name}: {d}"


def test_compare_tool_verdicts(tmp_path):
    """tools/compare_snapshots.py: identical files pass; a file written by the (.NET) reference carries no feature ids
    and is compared on pair sets and contact counts; a perturbation above the tolerance fails the run"""
    import sys
    name = sorted(CASES)[-1]
    src = os.path.join(GOLD, name + ".pstj")
    tool = os.path.join(ROOT, "tools", "compare_snapshots.py")
    out = subprocess.run([sys.executable, tool, src, src], capture_output=True, text=True)
    assert out.returncode == 0 and "discrete decisions identical in every frame" in out.stdout
    t = snapshot.load(src)
    t.producer = snapshot.PRODUCER_REFERENCE_DOTNET
    for f in t.frames:
        for c in f.contacts:
            c["feature"][:] = 0
    dyn = np.flatnonzero(~np.isinf(t.constants["mass"]))
    t.frames[-1].state["vel"][dyn[0], 2] += 1e-11
    ref = str(tmp_path / "ref.pstj")
    snapshot.save(ref, t)
    out = subprocess.run([sys.executable, tool, ref, src, "--rtol", "1e-9"], capture_output=True, text=True)
    assert out.returncode == 0 and "reference (.NET)" in out.stdout and "within rtol" in out.stdout
    out = subprocess.run([sys.executable, tool, ref, src, "--rtol", "1e-13"], capture_output=True, text=True)
    assert out.returncode == 1 and "ABOVE rtol" in out.stdout
