"""Mirror of the reference's ``neural_cbf`` package: systems + controllers (inference methods only)."""
