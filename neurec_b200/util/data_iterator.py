"""DataIterator: batches over parallel sequences, yielding python lists.

Mirror of the reference's util/data_iterator.py:158-210 (and its Sampler classes :15-109):
same constructor, same ``len()``, same batch contents.  Shuffling draws ONE
``np.random.permutation(n)`` per ``__iter__`` exactly like RandomSampler (:59), so under the same
numpy seed the batches are identical to the reference's.
"""
import numpy as np


class DataIterator(object):
    def __init__(self, *data, batch_size=1, shuffle=False, drop_last=False):
        for d in data:
            if len(d) != len(data[0]):
                raise ValueError("The length of the given data are not equal!")   # :139-141
        if not isinstance(batch_size, int) or isinstance(batch_size, bool) or batch_size <= 0:
            raise ValueError("batch_size should be a positive integeral value, "
                             "but got batch_size={}".format(batch_size))         # :84-88
        if not isinstance(drop_last, bool):
            raise ValueError("drop_last should be a boolean value, but got "
                             "drop_last={}".format(drop_last))                    # :89-91
        self.data = list(data)
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last

    def __len__(self):
        n = len(self.data[0]) if self.data else 0
        if self.drop_last:
            return n // self.batch_size
        return (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.data[0]) if self.data else 0
        order = np.random.permutation(n) if self.shuffle else np.arange(n)
        cols = [d if isinstance(d, np.ndarray) else np.asarray(d, dtype=object if _ragged(d) else None)
                for d in self.data]
        for off in range(0, n, self.batch_size):
            idx = order[off:off + self.batch_size]
            if len(idx) < self.batch_size and self.drop_last:
                return
            batch = [c[idx].tolist() for c in cols]
            yield batch[0] if len(batch) == 1 else batch           # :150-151 single array unwraps


def _ragged(seq):
    try:
        first = seq[0]
    except (IndexError, KeyError, TypeError):
        return False
    return isinstance(first, (list, tuple, np.ndarray))
