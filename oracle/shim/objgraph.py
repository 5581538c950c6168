"""Empty stand-in so the reference's env/workflow.py:14 imports in the build container (test infrastructure)."""
