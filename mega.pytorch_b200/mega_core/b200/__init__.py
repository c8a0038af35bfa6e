"""B200 kernel wrappers and the per-frame engine (host side above the C ABI)."""
