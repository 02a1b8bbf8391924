"""Deprecated argument names the reference still accepts (its ``deprecated_alias`` decorator, _version_utils.py:21-48):
the old keyword works with a FutureWarning; giving both spellings is a KeyError."""
import functools
import warnings


def renamed_arguments(**old_to_new):
    def decorate(fn):
        @functools.wraps(fn)
        def call(*args, **kwargs):
            for old, new in old_to_new.items():
                if old not in kwargs:
                    continue
                if new in kwargs:
                    raise KeyError(f"{fn.__name__} received both `{old}` (deprecated) and `{new}` (recommended)")
                warnings.warn(f"The argument `{old}` is deprecated for {fn.__name__}; use `{new}` instead.", FutureWarning)
                kwargs[new] = kwargs.pop(old)
            return fn(*args, **kwargs)
        return call
    return decorate
