def __getattr__(name):
    raise RuntimeError("matplotlib stub: rendering is out of scope (%s)" % name)
