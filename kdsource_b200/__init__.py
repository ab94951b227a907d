"""kdsource_b200: B200-native kernel-density hot path of KDSource.

Like the reference package, the top-level module is intentionally empty
(``python/kdsource/__init__.py.in:39-40``); use ``import kdsource_b200.api as kds``.
"""
