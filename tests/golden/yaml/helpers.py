"""Constraint helpers referenced from mixed_grammar.yaml through `source:`."""
OFFSET = 0.25


def penalty(count, step):
    if count == 1:
        return step * 10
    return step / 4
