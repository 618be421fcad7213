"""Shared definition of the seeded GCN golden cases (used by make_golden.py and the tests)."""
SMALL = dict(lstm_hidden=64, lm_dim=128, gc_dims=(64, 96), fc_dim=64, n_terms=24, logit_scale=0.05)

# tag -> (GCNConfig kwargs, model seed, n proteins, Lmin, Lmax)
GCN_CASES = {
    "small": (dict(**SMALL), 5, 6, 8, 90),
    "small_relu_bias": (dict(gc_activation="Relu", gc_bias=True, **SMALL), 6, 4, 8, 90),
    "mf": (dict(), 1234, 3, 40, 140),
    # smallest shape the tensor-core engine accepts (H % 64, E % 128, g % 128)
    "tc_small": (dict(lstm_hidden=64, lm_dim=128, gc_dims=(128, 128), fc_dim=64, n_terms=24, logit_scale=0.05), 8, 9, 1, 300),
}
THRESHOLD, GEN = 10.0, 2


def workload_seed(model_seed: int) -> int:
    return model_seed + 100
