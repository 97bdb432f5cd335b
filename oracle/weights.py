"""Seeded random-init weights (moved to flamed_tts_b200/synthetic.py so that bench.py's product arm
does not import from oracle/); re-exported here for the oracle and the tests."""
from flamed_tts_b200.synthetic import *  # noqa: F401,F403
from flamed_tts_b200.synthetic import (N_SYMBOLS, kaiser_sinc_filter, make_codec_decoder_state_dict,  # noqa: F401
                                       make_codec_encoder_state_dict, make_flamed_state_dict, sinusoid_table)
