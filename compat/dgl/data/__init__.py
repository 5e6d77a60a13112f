"""dgl.data.DGLDataset protocol (dataloader.py:8,15): has_cache() ? load() : (download(); process(); save())."""
import os

from . import utils  # noqa: F401


class DGLDataset(object):
    def __init__(self, name, url=None, raw_dir=None, save_dir=None, hash_key=(), force_reload=False, verbose=False, transform=None):
        self._name = name
        self._url = url
        self._force_reload = force_reload
        self._verbose = verbose
        self._hash_key = hash_key
        self._transform = transform
        self._raw_dir = raw_dir if raw_dir is not None else os.path.join(os.path.expanduser("~"), ".dgl")
        self._save_dir = save_dir if save_dir is not None else self._raw_dir
        self._load()

    # ---- hooks subclasses override -------------------------------------------------------------
    def download(self):
        pass

    def process(self):
        raise NotImplementedError

    def save(self):
        pass

    def load(self):
        pass

    def has_cache(self):
        return False

    def __getitem__(self, idx):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    # ---- driver ----------------------------------------------------------------------------------
    def _download(self):
        if self._url is not None and not os.path.exists(self.raw_path):
            os.makedirs(self.raw_dir, exist_ok=True)
            self.download()

    def _load(self):
        loaded = False
        if not self._force_reload and self.has_cache():
            try:
                self.load()
                loaded = True
                if self.verbose:
                    print("Done loading data from cached files.")
            except KeyboardInterrupt:
                raise
            except Exception as exc:        # a stale or foreign cache: rebuild, as DGL does
                if self.verbose:
                    print(f"Loading from cache failed ({exc}); re-processing.")
        if not loaded:
            self._download()
            self.process()
            self.save()
            if self.verbose:
                print("Done saving data into cached files.")

    # ---- properties ------------------------------------------------------------------------------
    @property
    def url(self):
        return self._url

    @property
    def name(self):
        return self._name

    @property
    def raw_dir(self):
        return self._raw_dir

    @property
    def raw_path(self):
        return os.path.join(self.raw_dir, self.name)

    @property
    def save_dir(self):
        return self._save_dir

    @property
    def save_path(self):
        return os.path.join(self.save_dir, self.name)

    @property
    def verbose(self):
        return self._verbose

    @property
    def hash(self):
        return str(abs(hash(self._hash_key)))[:8]
