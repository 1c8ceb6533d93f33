"""Haiku module paths of the reference's wave-function parameters -- the ORACLE's own table (test infrastructure).

The names are data of the reference, not of the product: ``hk.transform`` names a parameter ``<module path>:<param>`` (':'-joined
by the reference's tests/conftest.py:39-52 ``flatten_pytree``), and the paths below can be read off the arrays of the reference's
tests/test_wf/test_grad_psi.npz (conv-GNN test ansatz) and off the module nesting in wf/nn_wave_function.py (the root module),
wf/omni.py:13-211 (OmniNet: Backflow / Backflow_1 / Jastrow / nuclear_gnn_head), gnn/electron_gnn.py:243-625 (electron_gnn,
electron_gnn_layer[_i], electron_embedding, nuclei_embedding), gnn/update_features.py (the update-feature modules), wf/env.py and
wf/cusp.py.  The product keeps its own copy (deepqmc_b200/params.py); tests/test_oracle_goldens.py pins both against the reference's
recorded parameter table, so the checker does not import names from the thing it checks.
"""

ROOT = 'neural_network_wave_function/~/'
ENV = ROOT + 'exponential_envelopes'
CUSP = ROOT + 'electronic_cusp_asymptotic'
NUC_CUSP = ROOT + 'nuclear_cusp_asymptotic'
CONF = ROOT + 'conf_coeff'
OMNI = ROOT + 'omni_net/~/'
GNN = OMNI + 'electron_gnn/~/'
JASTROW = OMNI + 'Jastrow/~/mlp/'
HEAD = OMNI + 'nuclear_gnn_head/'
NUC_EMB = GNN + 'nuclei_embedding/'
# Backflow(multi_head=True): one net per spin (wf/omni.py:43-88); a second net "mlp_1" for backflow_transform = 'both' (:69-73)
BF_UP = OMNI + 'Backflow/~/mlp/linear_0'
BF_DN = OMNI + 'Backflow_1/~/mlp/linear_0'
BF_UP_ADD = OMNI + 'Backflow/~/mlp_1/linear_0'
BF_DN_ADD = OMNI + 'Backflow_1/~/mlp_1/linear_0'


def layer_prefix(l: int) -> str:
    """haiku numbers repeated modules: electron_gnn_layer, electron_gnn_layer_1, ..."""
    return GNN + ('electron_gnn_layer' if l == 0 else f'electron_gnn_layer_{l}') + '/~/'


def attn_prefix(l: int) -> str:
    return layer_prefix(l) + 'node_attention_electron_update_feature/'


def comb_prefix(l: int) -> str:
    return layer_prefix(l) + 'combined_node_attention_update_feature/'


def conv_prefix(l: int) -> str:
    return layer_prefix(l) + 'convolution_electron_update_feature/~single_edge_type_update/'
