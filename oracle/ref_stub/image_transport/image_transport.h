// image_transport/image_transport.h — STUB (oracle/_ref)
#pragma once
