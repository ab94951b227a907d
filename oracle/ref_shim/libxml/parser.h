/* Shim for <libxml/parser.h> (libxml2 is not installed here).  Only the
 * declarations KDS_open needs to compile; the harness never calls KDS_open,
 * it builds sources with KDS_create/PList_create/Metric_create/Geom_create.
 * TEST INFRASTRUCTURE (oracle/_ref build) -- not part of the product. */
#ifndef ORC_SHIM_LIBXML_PARSER_H
#define ORC_SHIM_LIBXML_PARSER_H
typedef unsigned char xmlChar;
typedef struct _xmlNode {
  const xmlChar *name;
  struct _xmlNode *children;
  struct _xmlNode *next;
} xmlNode;
typedef xmlNode *xmlNodePtr;
typedef struct _xmlDoc { int unused; } xmlDoc;
typedef xmlDoc *xmlDocPtr;
int xmlKeepBlanksDefault(int);
xmlDocPtr xmlReadFile(const char *, const char *, int);
xmlNodePtr xmlDocGetRootElement(xmlDocPtr);
xmlChar *xmlNodeGetContent(xmlNodePtr);
xmlChar *xmlGetProp(xmlNodePtr, const xmlChar *);
void xmlFreeDoc(xmlDocPtr);
void xmlCleanupParser(void);
#endif
