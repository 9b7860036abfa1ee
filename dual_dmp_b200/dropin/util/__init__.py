"""`util` shim: put <repo>/dual_dmp_b200/dropin (and <repo>) on PYTHONPATH and the reference drivers' `import util.*`
resolve to the B200 path without editing them."""
