"""Column hashing + Merkle tree of the linear-code commitments (csrc/hash.cuh) against hashlib:
BLAKE2s-256 / SHA-256 of to_bytes!(column) and the SHA-256 tree with ark-crypto-primitives' framing (ByteDigestConverter at
the leaf level, empty padding leaves), restated here with hashlib only.  RFC 7693 / FIPS 180-4 'abc' vectors pin hashlib
itself (tests/golden/external_kats.json).  CPU: host-emulated kernels; GPU: the CUDA library at the 2^20-coefficient shape."""
import hashlib
import struct

import numpy as np
import pytest

from oracle import orc, pyref
from tests import util


@pytest.fixture(scope="module")
def emul(pc, hostcheck_path):
    e = pc.Engine(0, lib_path=hostcheck_path)
    yield e
    e.close()


def ref_column_hashes(C, ext_mat_mont, hash_name):
    """leaves[j] = D(u64 n_rows || canonical LE bytes of column j): F::into_bigint through the C oracle, digest by hashlib"""
    n_rows, n_cols = ext_mat_mont.shape[0], ext_mat_mont.shape[1]
    canon = orc.field_unop("orc_fr_from_mont", C.id, ext_mat_mont.reshape(-1, 4)).reshape(n_rows, n_cols, 4)
    out = []
    for j in range(n_cols):
        h = hashlib.new(hash_name)
        h.update(struct.pack("<Q", n_rows) + np.ascontiguousarray(canon[:, j, :]).astype("<u8").tobytes())
        out.append(h.digest())
    return out


def ref_merkle(leaves):
    """heap-ordered inner nodes + root, ark-crypto-primitives framing"""
    P = 1 << max(1, (len(leaves) - 1).bit_length())
    padded = list(leaves) + [b""] * (P - len(leaves))
    conv = lambda d: struct.pack("<Q", len(d)) + d                                  # ByteDigestConverter: to_uncompressed_bytes!(Vec<u8>)
    level = [hashlib.sha256(conv(padded[2 * i]) + conv(padded[2 * i + 1])).digest() for i in range(P // 2)]
    levels = [level]
    while len(level) > 1:
        level = [hashlib.sha256(level[2 * i] + level[2 * i + 1]).digest() for i in range(len(level) // 2)]
        levels.append(level)
    nodes = [d for lv in reversed(levels) for d in lv]                               # root first, then level by level
    return nodes, levels[-1][0]


def _check(eng, cname, n_rows, n_cols, rho_inv, seed):
    from poly_commit_b200 import linear_codes
    C = pyref.Curve(cname)
    mat = util.rand_fr(cname, n_rows * n_cols, seed=seed, mont=True).reshape(n_rows, n_cols, 4)
    log_ext = max(1, (n_cols * rho_inv - 1).bit_length())
    ext = eng.ntt_batch(C.id, mat, log_ext)
    for hid, hname in ((0, "blake2s"), (1, "sha256")):
        exp_leaves = ref_column_hashes(C, ext, hname)
        got = eng.lincode_hash_columns(C.id, ext, hash=hid)
        assert [bytes(r) for r in got] == exp_leaves, (cname, hname)
        nodes, root = eng.merkle_tree(got)
        exp_nodes, exp_root = ref_merkle(exp_leaves)
        assert root.tobytes() == exp_root and [bytes(r) for r in nodes] == exp_nodes
        fused = eng.lincode_commit(C.id, mat, log_ext, hash=hid)
        assert fused["root"].tobytes() == exp_root and (fused["ext"] == ext).all()
        assert (fused["leaves"] == got).all() and (fused["nodes"] == nodes).all()
    # a leaf count that is not a power of two: padding leaves are empty
    odd = got[: max(2, (1 << log_ext) - 3)]
    nodes, root = eng.merkle_tree(odd)
    exp_nodes, exp_root = ref_merkle([bytes(r) for r in odd])
    assert root.tobytes() == exp_root and [bytes(r) for r in nodes] == exp_nodes
    # authentication paths recompute the root
    st = dict(leaves=got, nodes=eng.merkle_tree(got)[0])
    root = eng.merkle_tree(got)[1].tobytes()
    for i in (0, 1, (1 << log_ext) - 1, (1 << log_ext) // 3):
        p = linear_codes.merkle_path(st, i)
        conv = lambda d: struct.pack("<Q", len(d)) + d
        me, sib = got[i].tobytes(), p["leaf_sibling_hash"]
        cur = hashlib.sha256(conv(me) + conv(sib) if i % 2 == 0 else conv(sib) + conv(me)).digest()
        idx = i // 2
        for d in reversed(p["auth_path"]):
            cur = hashlib.sha256(cur + d if idx % 2 == 0 else d + cur).digest()
            idx //= 2
        assert cur == root


@pytest.mark.parametrize("cname,n_rows,n_cols", [("bls12_381", 5, 6), ("bn254", 8, 16), ("pallas", 1, 3), ("bls12_381", 2, 1)])
def test_column_hash_and_merkle_vs_hashlib(emul, cname, n_rows, n_cols):
    _check(emul, cname, n_rows, n_cols, 4, seed=300 + n_rows)


def test_ligero_commit_mirror(emul, pc):
    """linear_codes.commit: dimensions as ligero.rs:118-128, root and state from one device call"""
    from poly_commit_b200 import linear_codes
    cname = "bn254"
    C = pyref.Curve(cname)
    coeffs = util.rand_fr(cname, 300, seed=310, mont=True)
    comm, st = linear_codes.commit(emul, C.id, coeffs, sec_param=128, rho_inv=4)
    n_rows, n_cols, n_ext = comm["metadata"]
    assert n_rows * n_cols >= 300 and n_ext == 1 << max(0, (n_cols * 4 - 1).bit_length())
    leaves = ref_column_hashes(C, st["ext_mat"], "blake2s")
    assert comm["root"] == ref_merkle(leaves)[1]
    assert (st["ext_mat"] == linear_codes.compute_matrices(emul, C.id, coeffs, n_rows, n_cols, 4)[1]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("cname,n_rows,n_cols", [("bls12_381", 5, 6), ("bn254", 64, 100), ("pallas", 1, 3), ("bls12_381", 33, 700)])
def test_gpu_column_hash_and_merkle_vs_hashlib(gpu_engine, cname, n_rows, n_cols):
    _check(gpu_engine, cname, n_rows, n_cols, 4, seed=320 + n_rows)


@pytest.mark.gpu
def test_gpu_ligero_commit_2p20_shape(gpu_engine, pc):
    """the 2^20-coefficient Ligero shape (BLS12-381 Fr, rho_inv = 4): fused commit == separate calls; a sample of columns
    against hashlib; the tree against hashlib over the device's leaves"""
    from poly_commit_b200 import linear_codes
    eng, cname = gpu_engine, "bls12_381"
    C = pyref.Curve(cname)
    coeffs = util.rand_fr_fast(cname, 1 << 20, seed=330)
    comm, st = linear_codes.commit(eng, C.id, coeffs)
    n_rows, n_cols, n_ext = comm["metadata"]
    ext = eng.ntt_batch(C.id, st["mat"], n_ext.bit_length() - 1)
    assert (ext == st["ext_mat"]).all()
    cols = [0, 1, n_ext // 2 + 5, n_ext - 1]
    sample = np.ascontiguousarray(ext[:, cols, :])
    exp = ref_column_hashes(C, sample, "blake2s")
    assert [st["leaves"][j].tobytes() for j in cols] == exp
    nodes, root = ref_merkle([bytes(r) for r in st["leaves"]])
    assert comm["root"] == root and [bytes(r) for r in st["nodes"]] == nodes
