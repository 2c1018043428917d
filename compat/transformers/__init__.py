"""`transformers` namespace shim (the reference vendors pytorch_transformers 1.0.0 under this name, and imports
it as `transformers.pytorch_transformers`).  Only that subpackage is served here."""
