"""TEST INFRASTRUCTURE ONLY: empty stand-in so the reference module's top-level
``import matplotlib.pyplot`` succeeds; no plotting function is ever called by the harness."""
