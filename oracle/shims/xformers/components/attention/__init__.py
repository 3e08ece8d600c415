class AttentionMask:  # imported by the reference, never used on the path
    pass
