"""No-op stand-in so the reference's import-time matplotlib use works offline (golden generation only)."""
rcParams = {}
def use(*a, **k): pass
