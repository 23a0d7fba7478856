from scipy.stats import norm  # noqa: F401  (only imported by jaxpm/utils.py, off the hot path)
