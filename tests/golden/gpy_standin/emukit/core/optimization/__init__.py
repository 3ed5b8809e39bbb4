"""Names imported by the reference's causal_optimizer.py.  The L-BFGS refinement is outside the rebuilt path (the grid
argmax replaces it), so these exist only to let ``import src.utils_functions`` succeed."""
