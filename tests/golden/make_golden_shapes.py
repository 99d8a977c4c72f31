"""Golden vectors for the policy shapes the reference exposes on its command line beyond the defaults
(exp_runners/env_uitils.py:82-87: --encoder_hidden_sizes, --embedding_dim, --attention_type, --categorical_mlp_hidden_sizes):
the UNMODIFIED reference CommCategoricalMLPPolicy with narrower layers and with attention_type='dot'
(attention_module.py:38-49), evaluated like the default-shape fixtures (tests/golden/make_golden.py::run_policy_case).
Run in the build container (needs /root/reference):  python tests/golden/make_golden_shapes.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
import ref_harness as H  # noqa: E402


def main():
    ns = H.load_reference()
    P = H.scenario_params
    G.run_policy_case(ns, "narrow_c3", "pp", P("pp", 20, 2, 0.08, cap=4, loss=0.2), B=3, seed=71, loss_for_masks="zero_row",
                      encoder_hidden_sizes=(96,), embedding_dim=48, categorical_mlp_hidden_sizes=(64, 48, 16))
    G.run_policy_case(ns, "dot_c1", "pp", P("pp", 10, 1, 0.04, cap=2, loss=0.3), B=5, seed=72, loss_for_masks=None, attention_type="dot")
    G.run_policy_case(ns, "dot_narrow_c4", "co", P("co", 30, 2, 0.06, loss=0.1), B=2, seed=73, loss_for_masks=None,
                      attention_type="dot", encoder_hidden_sizes=(40,), embedding_dim=24, categorical_mlp_hidden_sizes=(32, 24, 8))


if __name__ == "__main__":
    main()
