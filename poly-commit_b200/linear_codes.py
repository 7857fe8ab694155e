"""Host mirror of the Ligero prover's matrix step (poly-commit/src/linear_codes) over the C ABI -- SURVEY.md section 8f
rank 4: the one place the reference exercises the NTT.

  calculate_t            linear_codes/utils.rs:156-184
  compute_dimensions     linear_codes/ligero.rs:118-128
  Matrix::new_from_flat  utils.rs:61-77            (row-major: entry[row][col] = flat[m * row + col])
  reed_solomon           linear_codes/utils.rs:112-127   -> one row of Engine.ntt_batch
  compute_matrices       linear_codes/mod.rs:118-138     -> Engine.ntt_batch over all rows (one launch up to 2^11 columns)
  b^T . M                linear_codes/mod.rs:? (open) via Matrix::row_mul, utils.rs:127-146  -> Engine.fr_row_mul

  commit steps 2-3       linear_codes/mod.rs:253-275: column hashes (FieldToBytesColHasher<F, Blake2s256>) and the Merkle tree
                         (LeafIdentityHasher + SHA-256 two-to-one, the configuration of the reference's tests and benches) run on
                         the device behind Engine.lincode_commit; `commit` below is LinearCodePCS::commit for one polynomial and
                         `merkle_path` is MerkleTree::generate_proof (used by generate_proof, linear_codes/mod.rs:553-558).
"""
import math

import numpy as np

FIELD_BITS = {0: 255, 1: 254, 2: 255}


def ceil_div(x, y):
    return (x + y - 1) // y


def calculate_t(field_bits, sec_param, distance, codeword_len):
    """linear_codes/utils.rs:156-184 (same f64 arithmetic); field_bits = F::MODULUS_BIT_SIZE"""
    residual = codeword_len / 2.0 ** field_bits
    arg = 2.0 ** (-sec_param) - residual
    if not arg > 0.0:
        raise ValueError("InvalidParameters: the field is not big enough")
    nom = math.log2(arg) - 1.0
    denom = math.log2(1.0 - 0.5 * distance[0] / distance[1])
    if denom == 0.0 or not math.isfinite(denom):
        raise ValueError("InvalidParameters: the distance is wrong")
    t = math.ceil(nom / denom)
    return t if t < codeword_len else codeword_len


def compute_dimensions(curve, sec_param, rho_inv, poly_len):
    """linear_codes/ligero.rs:118-128 with distance = (rho_inv - 1, rho_inv) (ligero.rs distance())"""
    t = calculate_t(FIELD_BITS[curve], sec_param, (rho_inv - 1, rho_inv), poly_len)
    root = math.ceil(math.sqrt(ceil_div(2 * poly_len, t)))
    n = 1 << max(0, (root - 1).bit_length())                 # 1 << log2(x): ark_std::log2 is ceil(log2 x)
    return n, ceil_div(poly_len, n)


def _domain_log(size):
    return max(0, (size - 1).bit_length())


def reed_solomon(eng, curve, msg, rho_inv):
    """linear_codes/utils.rs:112-127: msg (m, 4) -> evaluations over the smallest domain of size >= m * rho_inv"""
    msg = np.asarray(msg, dtype=np.uint64).reshape(1, -1, 4)
    return eng.ntt_batch(curve, msg, _domain_log(msg.shape[1] * rho_inv))[0]


def compute_matrices(eng, curve, coeffs, n_rows, n_cols, rho_inv):
    """linear_codes/mod.rs:118-138 -> (mat (n_rows, n_cols, 4), ext_mat (n_rows, domain, 4)); coeffs zero-padded to n_rows*n_cols"""
    coeffs = np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    if coeffs.shape[0] > n_rows * n_cols:
        raise ValueError("more coefficients than matrix entries")
    flat = np.zeros((n_rows * n_cols, 4), dtype=np.uint64)
    flat[:coeffs.shape[0]] = coeffs
    mat = flat.reshape(n_rows, n_cols, 4)
    return mat, eng.ntt_batch(curve, mat, _domain_log(n_cols * rho_inv))


BLAKE2S, SHA256 = 0, 1


def commit(eng, curve, coeffs, sec_param=128, rho_inv=4, hash=BLAKE2S):
    """LinearCodePCS::commit for one polynomial given as its coefficient vector (linear_codes/mod.rs:228-298) ->
    (commitment dict(metadata=(n_rows, n_cols, n_ext_cols), root=bytes), state dict(mat, ext_mat, leaves, nodes)).
    Row encoding, column hashing and the tree run in ONE device-resident call."""
    coeffs = np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    n_rows, n_cols = compute_dimensions(curve, sec_param, rho_inv, coeffs.shape[0])
    flat = np.zeros((n_rows * n_cols, 4), dtype=np.uint64)
    flat[:coeffs.shape[0]] = coeffs
    mat = flat.reshape(n_rows, n_cols, 4)
    log_ext = _domain_log(n_cols * rho_inv)
    r = eng.lincode_commit(curve, mat, log_ext, hash=hash)
    n_ext = 1 << log_ext
    return (dict(metadata=(n_rows, n_cols, n_ext), root=r["root"].tobytes()),
            dict(mat=mat, ext_mat=r["ext"], leaves=r["leaves"], nodes=r["nodes"]))


def merkle_path(state, index):
    """MerkleTree::generate_proof(index) over the state's tree -> dict(leaf_sibling_hash, auth_path (root's children first, as
    ark-crypto-primitives stores it), leaf_index).  The padding leaves are empty (Vec::default())."""
    leaves, nodes = state["leaves"], state["nodes"]
    P = nodes.shape[0] + 1
    sib = index ^ 1
    leaf_sibling = leaves[sib].tobytes() if sib < leaves.shape[0] else b""
    node = P // 2 - 1 + index // 2                    # heap index of the leaf pair's parent
    path = []
    while node > 0:
        sibling = node + 1 if node % 2 == 1 else node - 1
        path.append(nodes[sibling].tobytes())
        node = (node - 1) // 2
    return dict(leaf_sibling_hash=leaf_sibling, auth_path=path[::-1], leaf_index=index)
