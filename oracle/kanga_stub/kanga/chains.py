class ChainArray:
    def __init__(self, vals):
        self.vals = vals
