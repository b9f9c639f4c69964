"""Thin command-line callers with the reference's flag surface (2d_u1_cluster/compare_hmc.py,
compare_fthmc.py, conf_gen.py): argument parsing and timing prints only — the work is in the
sampler classes."""
