"""Stand-in for env/workflow.py:15 (test infrastructure)."""


def mem_top(*a, **k):
    return ""
