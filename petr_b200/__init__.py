"""petr_b200: B200-native (sm_100a) implementation of the PETR / PETRv2 detection-head hot path
behind the reference's mmdet3d plugin API (include/petr_b200.h)."""
from .registry import (ATTENTION, BBOX_CODERS, FEEDFORWARD_NETWORK, HEADS, POSITIONAL_ENCODING,
                       NECKS, TRANSFORMER, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE, build_head,
                       build_neck, build_transformer, register_into_mm)
from .modules import (FFN, LearnedPositionalEncoding3D, MultiheadAttention, NMSFreeCoder, PETRHead,
                      PETRMultiheadAttention, PETRTransformer, PETRTransformerDecoder,
                      PETRTransformerDecoderLayer, PETRTransformerEncoder, PETRv2Head,
                      SinePositionalEncoding, SinePositionalEncoding3D, NormedLinear, DepthrHead, set_precision)
from .necks import CPFPN
from .configs import WORKLOADS, petr_head_cfg, petrv2_head_cfg

register_into_mm()

__all__ = [n for n in dir() if not n.startswith("_")]
