from .get_model import Args, model_dict, get_model_kwargs, calculate_input_channels  # noqa: F401
from .popcorn import POPCORN  # noqa: F401
