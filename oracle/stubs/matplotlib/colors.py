"""matplotlib.colors stand-in (never executed on the render path)."""
