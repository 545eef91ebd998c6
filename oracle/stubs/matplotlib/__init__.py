"""Stand-in for `matplotlib` (absent offline; imported, never executed, on the SceneDreamer render path)."""


def use(*a, **k):
    pass
