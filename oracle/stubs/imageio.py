"""Stand-in for `imageio` (absent offline).  The reference only uses `get_writer(...).append_data/close`
(imaginaire/generators/scenedreamer.py:558,628-629) to write an .mp4 next to the PNG frames; the
checker does not need the video, so frames are dropped."""


class _NullWriter:
    def __init__(self, path, **kw):
        self.path, self.frames = path, 0

    def append_data(self, frame):
        self.frames += 1

    def close(self):
        pass


def get_writer(path, **kw):
    return _NullWriter(path, **kw)


def imread(*a, **k):
    raise RuntimeError('imageio stub: imread is not available offline')


imwrite = imsave = imread
