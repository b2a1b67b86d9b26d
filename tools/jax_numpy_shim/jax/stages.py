class Lowered:  # placeholder for type annotations
    pass
