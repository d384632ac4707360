"""Helpers either side of the hot path: ``helpers`` / ``initialization`` (shape and init utilities of the models),
``datasplit`` (context / target split on the device), ``gp`` (synthetic GP tasks on the device), ``checkpoint`` (upstream's
run-directory layout) and ``train`` (``train_models`` / ``eval_loglike`` with upstream's signatures, skorch-free)."""
