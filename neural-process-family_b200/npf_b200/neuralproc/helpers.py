"""Reshape helpers for the (n_z_samples, batch) axes -- same contracts as upstream npf/neuralproc/helpers.py."""


def collapse_z_samples_batch(t):
    n_z, B, *rest = t.shape
    return t.contiguous().view(n_z * B, *rest)


def extract_z_samples_batch(t, n_z_samples, batch_size):
    return t.view(n_z_samples, batch_size, *t.shape[1:])


def replicate_z_samples(t, n_z_samples):
    return t.unsqueeze(0).expand(n_z_samples, *t.shape)


def pool_and_replicate_middle(t):
    first, *middle, last = t.shape
    pooled = t.reshape(first, -1, last).mean(1, keepdim=True)
    return pooled.view(first, *([1] * len(middle)), last).expand(first, *middle, last)
