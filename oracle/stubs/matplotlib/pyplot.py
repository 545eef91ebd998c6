"""matplotlib.pyplot stand-in (never executed on the render path)."""
