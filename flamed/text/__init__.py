"""Text front-end boundary.

Out of scope for the B200 hot path: the reference's cleaners
need g2p_en / inflect / unidecode and a lexicon blob that are not available offline.
`text_to_sequence` here handles the form `Flamed._preprocess_english` produces - ARPAbet
phones in curly braces - plus plain symbols, which is what the synthesis scripts need once a
lexicon is supplied.
"""
import re

from .symbols import symbol_to_id, symbols  # noqa: F401

_braces = re.compile(r"\{([^}]*)\}")


def text_to_sequence(text, cleaner_names=None):
    """'{HH AH0 L OW1} ,' -> [ids].  Tokens inside braces are ARPAbet ('@'-prefixed in the
    table); characters outside braces map to themselves; unknown tokens are dropped
    (reference: flamed/text/__init__.py:15-75)."""
    seq, pos = [], 0
    for m in _braces.finditer(text):
        seq += [symbol_to_id[c] for c in text[pos:m.start()] if c in symbol_to_id and c not in "_~"]
        seq += [symbol_to_id["@" + t] for t in m.group(1).split() if "@" + t in symbol_to_id]
        pos = m.end()
    seq += [symbol_to_id[c] for c in text[pos:] if c in symbol_to_id and c not in "_~"]
    return seq
