from .preprocessor import Preprocessor  # noqa: F401
from .batch_preprocessor import (  # noqa: F401
    BatchPreprocessor, DiscreteDqnBatchPreprocessor, PolicyNetworkBatchPreprocessor, batch_to_device)
