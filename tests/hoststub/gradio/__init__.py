"""gradio stand-in (TEST INFRASTRUCTURE): Info/Warning are no-ops, every widget is an inert context manager."""
messages = []


def Info(msg, *a, **k):
    messages.append(("info", msg))


def Warning(msg, *a, **k):  # noqa: A001
    messages.append(("warning", msg))


class _Widget:
    def __init__(self, *a, **k):
        self.value = k.get("value")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __getattr__(self, name):
        return lambda *a, **k: None


def __getattr__(name):
    return _Widget
