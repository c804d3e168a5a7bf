from . import type_check  # noqa
