"""Prints, for the named golden fixtures (default: all), the max-rel error of loc / scale / per-task loss against the
reference's golden vectors in each precision mode -- the evidence behind the tolerances written in tests/.
    python profiles/microbench/fixture_errors.py [substring ...]"""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "neural-process-family_b200")]
from _cfg import build_model, loss_for  # noqa: E402
from _util import fixture_names, load_fixture, rel_err  # noqa: E402
import npf_b200  # noqa: E402

subs = sys.argv[1:]
for name in fixture_names():
    if subs and not any(s in name for s in subs):
        continue
    fx = load_fixture(name)
    for prec in ("fp32", "bf16x3", "bf16"):
        npf_b200.set_precision(prec)
        try:
            model = build_model(fx["cfg"])
            model.load_state_dict(fx["state_dict"])
            model.cuda()
            for case in fx["cases"]:
                model.load_state_dict(fx["state_dict"])
                model.train(case["training"])
                if "eps" in case:
                    model._eps_override = case["eps"].cuda()
                inp = {k: v.cuda() for k, v in case["inputs"].items()}
                crit = loss_for(case["loss_name"])
                crit.train(case["training"])
                out = model(inp["X_cntxt"], inp["Y_cntxt"], inp["X_trgt"], inp["Y_trgt"])
                per_task = crit(out, inp["Y_trgt"])
                d = out[0].base_dist
                msg = f"{name}/{case['name']} [{prec}] loc {rel_err(d.loc, case['loc']):.2e} scale " \
                      f"{rel_err(d.scale, case['scale']):.2e} loss {rel_err(per_task, case['loss_per_task']):.2e}"
                if "q_loc" in case:
                    msg += f" q_loc {rel_err(out[2].base_dist.loc, case['q_loc']):.2e} q_scale {rel_err(out[2].base_dist.scale, case['q_scale']):.2e}"
                if case["training"]:
                    per_task.mean(0).backward()
                    torch.cuda.synchronize()
                    msg += " bwd ok"
                print(msg, flush=True)
        except Exception:
            print(f"{name} [{prec}] FAILED", flush=True)
            traceback.print_exc()
npf_b200.set_precision("fp32")
