"""Drop-in mirror of the reference's `utils` package (hot-path subset)."""
