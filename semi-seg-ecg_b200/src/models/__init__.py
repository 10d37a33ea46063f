"""Drop-in mirror of the reference's `models` package (src/models/ in the reference)."""
