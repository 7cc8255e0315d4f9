"""Host-side mirror of the reference's operator / plan interface (see executor.py, plan.py)."""
