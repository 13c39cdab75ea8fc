"""Drop-in replacement of the reference's `src` package for the stacked-hourglass hot path
(models / loss / utils / runner), backed by hand-written sm_100a kernels (hgb200)."""
