class Logexp:
    """Positive-constraint marker (the transform itself only matters to optimize(), which is absent here)."""
