"""Config C0 (BASELINE.json configs[0]): the shape of the reference's examples/wordcount.py:17-26
(map -> flat_map -> count_final -> StdOutSink) on a TestingSource."""
import re

import bytewax_b200.operators as op
from bytewax_b200.connectors.stdio import StdOutSink
from bytewax_b200.dataflow import Dataflow
from bytewax_b200.testing import TestingSource

LINES = [
    "To be, or not to be, that is the question:",
    "Whether 'tis nobler in the mind to suffer",
    "The slings and arrows of outrageous fortune,",
    "Or to take arms against a sea of troubles",
    "And by opposing end them.",
]

flow = Dataflow("wordcount_eg")
inp = op.input("inp", flow, TestingSource(LINES))
lower = op.map("lowercase_words", inp, str.lower)
tokens = op.flat_map("tokenize_input", lower, lambda line: re.findall(r'[^\s!,.?":;0-9]+', line))
counts = op.count_final("count", tokens, lambda word: word)
op.output("out", counts, StdOutSink())
