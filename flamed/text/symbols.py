"""Phoneme/character vocabulary of the Flamed-TTS text front-end.

The ID of a symbol is its position in `symbols`; the phoneme embedding table of every
checkpoint (`prior_generator.encoder.src_word_emb.weight`, 361 rows = 360 symbols + 1) is
indexed by it, so the table is part of the checkpoint contract and is shipped as data
(symbol_table.json: pad, '-', punctuation, letters, '@'-prefixed ARPAbet, '@'-prefixed
pinyin initials/finals, '@sp' '@spn' '@sil'; reference: flamed/text/symbols.py:21-29).
"""
import json
import os

with open(os.path.join(os.path.dirname(__file__), "symbol_table.json")) as _f:
    symbols = json.load(_f)

symbol_to_id = {s: i for i, s in enumerate(symbols)}
