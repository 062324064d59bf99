"""Model kwargs shared by make_attribution_golden.py and the tests."""
MODELS = {
    'plain': dict(dim_input=13, dim_output=1, k=32, num_layers=3,
                  edge_attention=True, node_attention=True, residual=True,
                  normalize=True, tanh=True, graphnorm=False),
    'gn3': dict(dim_input=13, dim_output=3, k=16, num_layers=2,
                edge_attention=True, node_attention=True, residual=True,
                normalize=True, tanh=True, graphnorm=True,
                model_task='multi_regression'),
    'gn1': dict(dim_input=13, dim_output=1, k=16, num_layers=2,
                edge_attention=True, node_attention=False, residual=True,
                normalize=False, tanh=False, graphnorm=True),
}
