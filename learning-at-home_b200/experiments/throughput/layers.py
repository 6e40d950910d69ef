"""re-export of the expert blocks under the reference's module path (experiments/throughput/layers.py)"""
from ...models.layers import FeedforwardBlock, TransformerEncoderLayer, name_to_block, name_to_input  # noqa: F401
