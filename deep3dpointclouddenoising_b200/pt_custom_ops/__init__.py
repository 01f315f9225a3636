"""Drop-in for the reference package u_net_arch/pt_custom_ops: `_ext` (the five compiled functions) and
`pt_utils` (autograd Functions + point-op modules), backed by libd3d_b200.so."""
