"""minimal stand-in (label table only); see ../README.md"""
