"""jyutvoice_b200: B200-native (sm_100a) text front + prompt encoder + CFM + HiFT hot path behind the reference's Python API."""
from .flow_matching import CausalConditionalCFM, CausalConditionalDecoder  # noqa: F401
from .hifigan import HiFTGenerator, ConvRNNF0Predictor  # noqa: F401
from .text import TextEncoder, DurationPredictor, length_regulate  # noqa: F401
from .tts import JyutVoiceTTS  # noqa: F401
from .flow_encoder import FlowEncoder, UpsampleConformerEncoder  # noqa: F401
from .checkpoint import load_checkpoint, load_hift, load_pretrain, split_flow_checkpoint, write_wav  # noqa: F401

__all__ = ["CausalConditionalCFM", "CausalConditionalDecoder", "HiFTGenerator", "ConvRNNF0Predictor", "TextEncoder",
           "DurationPredictor", "length_regulate", "JyutVoiceTTS", "FlowEncoder", "UpsampleConformerEncoder", "load_checkpoint", "load_hift", "load_pretrain",
           "split_flow_checkpoint", "write_wav"]
