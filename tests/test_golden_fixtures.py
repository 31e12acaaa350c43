"""tests/golden/ fixtures (written by scripts/make_golden.py).

* reference_kat.json: the reference's own known-answer vectors, transcribed from the cited test lines.
* oracle_vectors.npz: committed oracle outputs on seeded inputs.

`not gpu`: the oracle reproduces both.  `gpu`: the CUDA path, through the C ABI, reproduces both bit for bit
(trained centroids: bit-identical on one GPU with the ordered update, asserted at <= 1e-4 relative as north_star
states it).
"""
import json
import os

import numpy as np
import pytest

F = np.float32
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLD, "reference_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(GOLD, "oracle_vectors.npz"))


# ---------------------------------------------------------------------------------------------------------
# CPU: oracle against the fixtures
# ---------------------------------------------------------------------------------------------------------
def test_oracle_reference_kat(oracle, kat):
    p = kat["pq.rs:378-407"]
    q, x = np.array(p["test_quantizers"], F), np.array(p["test_vectors"], F)
    codes = np.array(p["test_quantizations"], np.uint64)
    assert np.array_equal(oracle.quantize_batch(q, None, x, np.uint64), codes)
    assert np.array_equal(oracle.reconstruct_batch(q, None, codes), np.array(p["test_reconstructions"], F))
    for v, c in zip(x, codes):
        assert np.array_equal(oracle.quantize_vector(q, None, v, np.uint64), c)
    a = kat["kmeans.rs:380-400"]
    assert oracle.cluster_assignments(np.array(a["centroids"], F), np.array(a["instances"], F)).tolist() == a["assignments"]
    u = kat["kmeans.rs:402-435"]
    got = oracle.update_centroids(np.array(u["centroids_before"], F), np.array(u["instances"], F),
                                  np.array(u["assignments"], np.uint64))
    assert np.array_equal(got, np.array(u["centroids_after"], F))
    m = kat["kmeans.rs:504-519"]
    mse = oracle.mean_squared_error(np.array(m["centroids"], F), np.array(m["instances"], F),
                                    np.array(m["assignments"], np.uint64))
    assert mse == F(m["mse_numerator"]) / F(m["mse_denominator"])
    d = kat["linalg.rs:298-313"]
    assert np.array_equal(oracle.sqdist_vec(np.array(d["a1"], F), np.array(d["b"], F)), np.array(d["sqdist_a1_b"], F))
    assert np.array_equal(oracle.sqdist_batch(np.array(d["a2"], F), np.array(d["b"], F)),
                          np.array(d["sqdist_a2_b"], F))


def test_bucket_eigenvalues_kat(kat):  # opq.rs:303-328 (host-side helper of the OPQ "next" row)
    from reductive_b200.opq import bucket_eigenvalues

    b = kat["opq.rs:303-328"]
    assert bucket_eigenvalues(b["eigenvalues_1"], b["n_buckets_1"]) == b["buckets_1"]
    assert bucket_eigenvalues(b["eigenvalues_2"], b["n_buckets_2"]) == b["buckets_2"]
    with pytest.raises(AssertionError):
        bucket_eigenvalues(b["uneven_panics"]["eigenvalues"], b["uneven_panics"]["n_buckets"])


def test_oracle_reproduces_committed_vectors(oracle, oracle_scalar, vec):
    for o in (oracle, oracle_scalar):
        for tag in "abc":
            q, x = vec[f"{tag}_q"], vec[f"{tag}_x"]
            codes = o.quantize_batch(q, None, x, np.uint8)
            assert np.array_equal(codes, vec[f"{tag}_codes"])
            assert np.array_equal(o.reconstruct_batch(q, None, codes), vec[f"{tag}_recon"])
            for i in range(16):
                assert np.array_equal(o.quantize_vector(q, None, x[i], np.uint8), vec[f"{tag}_vec_codes"][i])
        assert np.array_equal(o.quantize_batch(vec["a_q"], None, vec["tie_x"], np.uint8), vec["tie_codes"])
        assert np.array_equal(o.quantize_batch(vec["a_q"], vec["p_R"], vec["a_x"], np.uint8), vec["p_codes"])
        assert np.array_equal(o.reconstruct_batch(vec["a_q"], vec["p_R"], vec["p_codes"]), vec["p_recon"])
    cq, loss = oracle.train_pq(vec["k_x"], 4, 5, 4, 1, vec["k_init"])
    assert np.array_equal(cq, vec["k_q"])
    assert np.array_equal(np.asarray(loss, F), vec["k_loss"])


# ---------------------------------------------------------------------------------------------------------
# GPU: the CUDA path against the same committed bytes
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["exact", "auto"])
def test_cuda_reproduces_committed_vectors(vec, algo):
    import reductive_b200 as rb

    rb.set_encode_algo(rb.ENCODE_EXACT if algo == "exact" else rb.ENCODE_AUTO)
    try:
        for tag in "abc":
            pq = rb.Pq(None, vec[f"{tag}_q"])
            x = vec[f"{tag}_x"]
            codes = pq.quantize_batch(x, np.uint8)
            assert np.array_equal(codes, vec[f"{tag}_codes"]), tag
            assert np.array_equal(pq.reconstruct_batch(codes), vec[f"{tag}_recon"]), tag
            for i in range(16):
                assert np.array_equal(pq.quantize_vector(x[i], np.uint8), vec[f"{tag}_vec_codes"][i])
        pq = rb.Pq(None, vec["a_q"])
        assert np.array_equal(pq.quantize_batch(vec["tie_x"], np.uint8), vec["tie_codes"])
        pp = rb.Pq(vec["p_R"], vec["a_q"])
        assert np.array_equal(pp.quantize_batch(vec["a_x"], np.uint8), vec["p_codes"])
        rec = pp.reconstruct_batch(vec["p_codes"])
        err = np.linalg.norm(rec - vec["p_recon"]) / np.linalg.norm(vec["p_recon"])
        assert err <= 1e-5, err  # north_star: within 1e-5 relative after the rotation GEMM
    finally:
        rb.set_encode_algo(rb.ENCODE_AUTO)


@pytest.mark.gpu
def test_cuda_training_reproduces_committed_centroids(vec):
    import reductive_b200 as rb

    pq, loss = rb.Pq.train_pq_using(4, 5, 4, 1, vec["k_x"], np.random.default_rng(0),
                                    initial_centroids=vec["k_init"], return_loss=True)
    got = pq.subquantizers()
    rel = np.linalg.norm(got - vec["k_q"]) / np.linalg.norm(vec["k_q"])
    assert rel <= 1e-4, rel  # north_star: trained centroids within 1e-4 relative from identical initial centroids
    assert np.allclose(loss, vec["k_loss"], rtol=1e-3)


@pytest.mark.gpu
def test_cuda_reference_kat(kat):
    import reductive_b200 as rb

    p = kat["pq.rs:378-407"]
    pq = rb.Pq(None, np.array(p["test_quantizers"], F))
    codes = np.array(p["test_quantizations"], np.uint64)
    assert np.array_equal(pq.quantize_batch(np.array(p["test_vectors"], F), np.uint64), codes)
    assert np.array_equal(pq.reconstruct_batch(codes), np.array(p["test_reconstructions"], F))
    assert pq.quantized_len() == 2 and pq.reconstructed_len() == 6  # pq.rs:463-469
