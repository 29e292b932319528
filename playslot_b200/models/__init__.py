"""
Drop-in mirror of the reference's `src/models` package (same sub-packages, class names and exports,
src/models/__init__.py:5-11).
"""
from .EncodersDecoders.encoders import get_encoder
from .EncodersDecoders.decoders import get_decoder
from .BlocksUtils.attention import SlotAttention, MultiHeadSelfAttention, TransformerBlock
from .BlocksUtils.initializers import get_initializer
from .BlocksUtils.model_blocks import SoftPositionEmbed
from .BlocksUtils.model_utils import freeze_params
from .BlocksUtils.transition_models import get_transition_module
