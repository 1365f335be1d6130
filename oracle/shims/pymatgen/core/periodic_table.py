class Element:
    pass
