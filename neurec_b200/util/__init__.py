"""Host-side mirror of the reference's `util` package (TF-free)."""
from .configurator import Configurator
from .data_iterator import DataIterator
from .logger import Logger
from .tool import (csr_to_user_dict, csr_to_user_dict_bytime, pad_sequences, randint_choice,
                   timer, typeassert)


def batch_randint_choice(*args, **kwargs):
    from .random_choice import batch_randint_choice as f
    return f(*args, **kwargs)
